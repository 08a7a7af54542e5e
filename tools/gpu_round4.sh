#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --no-header -s -p no:cacheprovider > gpurun_out/test_gpu_conv.log 2>&1
echo "conv exit $?" >> gpurun_out/summary.txt; tail -n 8 gpurun_out/test_gpu_conv.log
timeout 300 python tools/conv_cases.py > gpurun_out/conv_cases_cur.log 2>&1; cat gpurun_out/conv_cases_cur.log
for c in c3x3_80 c3x3_320; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 1 -c 1 -f -o gpurun_out/ncu2_$c python tools/conv_cases.py $c 2 > gpurun_out/ncu2_$c.log 2>&1
  echo "ncu $c exit $?"
done
cat gpurun_out/summary.txt
