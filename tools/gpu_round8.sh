#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --no-header -s -p no:cacheprovider > gpurun_out/test_gpu_conv.log 2>&1
echo "conv exit $?" >> gpurun_out/summary.txt; tail -n 3 gpurun_out/test_gpu_conv.log
timeout 300 python tools/conv_cases.py > gpurun_out/conv_cases_cur.log 2>&1; cat gpurun_out/conv_cases_cur.log
timeout 600 python tools/profile_layers.py 64 gpurun_out/layers_bs64_cur.csv > gpurun_out/profile_cur.log 2>&1; echo "profile exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/profile_cur.log
cat gpurun_out/summary.txt
