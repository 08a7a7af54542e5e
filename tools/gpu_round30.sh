#!/bin/bash
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80
for pr in 0 3 19 11 27 16 8; do echo "== probe $pr"; MDB_CONV_PROBE=$pr python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["; done
