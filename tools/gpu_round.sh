#!/bin/bash
# Runs on the GPU box: full GPU test-suite, smoke, bench.  Each step in its own process.
mkdir -p gpurun_out
for f in test_gpu_letterbox test_gpu_nms test_gpu_conv test_gpu_e2e; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -x --no-header -s -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -n 12 gpurun_out/$f.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
cat gpurun_out/summary.txt
