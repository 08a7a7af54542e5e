#!/bin/bash
mkdir -p gpurun_out
MDB_CONV_DEBUG=1 timeout 300 python tools/conv_cases.py c1x1_160,c3x3_80,c3x3_320,stem 2 > gpurun_out/conv_dbg.log 2>&1; grep -E "dbg" gpurun_out/conv_dbg.log | awk '!seen[substr($0,1,70)]++' | cut -c1-330
for c in c1x1_160 c3x3_320; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 1 -c 1 -f -o gpurun_out/ncu3_$c python tools/conv_cases.py $c 2 > gpurun_out/ncu3_$c.log 2>&1
  echo "ncu $c exit $?"
done
