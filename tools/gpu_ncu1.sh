#!/bin/bash
mkdir -p gpurun_out
python tools/conv_cases.py > gpurun_out/conv_cases.log 2>&1; cat gpurun_out/conv_cases.log
for c in c1x1_160 c3x3_80 c3x3_320; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 1 -c 1 -f -o gpurun_out/ncu_$c python tools/conv_cases.py $c 2 > gpurun_out/ncu_$c.log 2>&1
  echo "ncu $c exit $?"; tail -2 gpurun_out/ncu_$c.log
done
ls -la gpurun_out
