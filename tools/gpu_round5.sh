#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --no-header -s -p no:cacheprovider > gpurun_out/test_gpu_conv.log 2>&1
echo "conv exit $?" >> gpurun_out/summary.txt; tail -n 6 gpurun_out/test_gpu_conv.log
timeout 300 python tools/conv_cases.py > gpurun_out/conv_cases_cur.log 2>&1; cat gpurun_out/conv_cases_cur.log
MDB_CONV_DEBUG=1 timeout 300 python tools/conv_cases.py c1x1_160,c3x3_80,c3x3_320,stem 2 2>&1 | grep -v "^\[conv dbg\].*" | head -0
MDB_CONV_DEBUG=1 timeout 300 python tools/conv_cases.py c1x1_160,c3x3_80,c3x3_320,stem 1 > gpurun_out/conv_dbg.log 2>&1; grep -E "dbg|ms" gpurun_out/conv_dbg.log | awk 'NR%2==1' | cut -c1-400
cat gpurun_out/summary.txt
