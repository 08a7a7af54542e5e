#!/bin/bash
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_160_80,b64_c3x3_80,b64_c1x1_320
echo "== baseline"; python tools/conv_cases.py $C 5 2>&1 | grep -v "^\[" 
echo "== debug"; MDB_CONV_DEBUG=1 python tools/conv_cases.py $C 2 2>&1 | tail -12
for pr in 1 2 3 4 5; do echo "== probe $pr"; MDB_CONV_PROBE=$pr timeout 300 python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["; done
echo "== halo80"; MDB_CONV_HALO_MIN=2 python tools/conv_cases.py b64_c3x3_80 5
