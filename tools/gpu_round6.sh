#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --no-header -s -p no:cacheprovider > gpurun_out/test_gpu_conv.log 2>&1
echo "conv exit $?" >> gpurun_out/summary.txt; tail -n 4 gpurun_out/test_gpu_conv.log
timeout 300 python tools/conv_cases.py > gpurun_out/conv_cases_cur.log 2>&1; cat gpurun_out/conv_cases_cur.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -s -p no:cacheprovider > gpurun_out/test_gpu_e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt; grep -E "e2e\]|passed|failed|Error" gpurun_out/test_gpu_e2e.log | cut -c1-900
timeout 600 python tools/profile_layers.py 64 gpurun_out/layers_bs64_cur.csv > gpurun_out/profile_cur.log 2>&1; echo "profile exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/profile_cur.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench_cur.json | cut -c1-1500
cat gpurun_out/summary.txt
