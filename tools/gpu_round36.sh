#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 1 -c 1 -f -o gpurun_out/ncu_c1x1_160 python tools/conv_cases.py b64_c1x1_160 2 > gpurun_out/ncu36a.log 2>&1; echo "exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 1 -c 1 -f -o gpurun_out/ncu_c1x1_80 python tools/conv_cases.py b64_c1x1_80 2 > gpurun_out/ncu36b.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
