#!/bin/bash
C=b64_c1x1_80,b64_c1x1_160,b64_c3x3_80
echo "== debug"; MDB_CONV_DEBUG=1 python tools/conv_cases.py $C 2 2>&1 | grep -v "^\[conv dbg\] grid" | awk 'NR%6==1||NR%6==0||/^b64/'
echo "== debug skeleton"; MDB_CONV_PROBE=7 MDB_CONV_DEBUG=1 python tools/conv_cases.py $C 2 2>&1 | tail -4
MDB_CONV_PROBE=7 MDB_CONV_DEBUG=1 python tools/conv_cases.py b64_c1x1_80 2 2>&1 | tail -3
MDB_CONV_PROBE=7 MDB_CONV_DEBUG=1 python tools/conv_cases.py b64_c1x1_160 2 2>&1 | tail -3
