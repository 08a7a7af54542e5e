#!/bin/bash
# Runs on the GPU box (via gpurun): each GPU test file in its own process (a device trap poisons a context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for f in test_gpu_letterbox test_gpu_nms test_gpu_conv; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -x --no-header -s -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -n 25 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
