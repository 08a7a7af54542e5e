#!/bin/bash
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80,c3x3_160,c3x3_320,c1x1_640,s2_80_160
for v in "" sl32 sl200 ew20 ew20sl; do echo "== variant '$v'"; MDB_LIB_VARIANT=$v python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["; done
echo "== skeleton default / sl200"; MDB_CONV_PROBE=7 python tools/conv_cases.py b64_c3x3_80,b64_c1x1_80 5;  MDB_LIB_VARIANT=sl200 MDB_CONV_PROBE=7 python tools/conv_cases.py b64_c3x3_80,b64_c1x1_80 5
