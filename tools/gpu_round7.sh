#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_conv test_gpu_nms test_gpu_e2e; do
timeout 900 python -m pytest tests/$f.py -m gpu -q --no-header -s -p no:cacheprovider > gpurun_out/$f.log 2>&1
echo "$f exit $?" >> gpurun_out/summary.txt; grep -E "e2e\]|passed|failed|Error" gpurun_out/$f.log | cut -c1-700
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -1 gpurun_out/smoke.log
timeout 300 python tools/conv_cases.py > gpurun_out/conv_cases_cur.log 2>&1; cat gpurun_out/conv_cases_cur.log
timeout 600 python tools/profile_layers.py 64 gpurun_out/layers_bs64_cur.csv > gpurun_out/profile_cur.log 2>&1; echo "profile exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/profile_cur.log
cat gpurun_out/summary.txt
