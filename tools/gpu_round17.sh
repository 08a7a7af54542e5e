#!/bin/bash
mkdir -p gpurun_out
for probe in 0 1 2; do
echo "== probe $probe (halo off)"; MDB_CONV_HALO=0 MDB_CONV_PROBE=$probe timeout 300 python tools/conv_cases.py c3x3_160,c3x3_320,c3x3_480,c1x1_320,c1x1_640 2>&1
done
