#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -2
for cps in 1 2 3; do
echo "== MDB_CONV_CPS=$cps"
MDB_CONV_CPS=$cps timeout 300 python tools/conv_cases.py 2>&1 | tee gpurun_out/conv_cases_cps$cps.log
done
