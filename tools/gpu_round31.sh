#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_conv.py -q -x 2>&1 | tail -3
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80,c3x3_160,c3x3_320,c1x1_640,s2_80_160,stem
python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
echo "== probe 3"; MDB_CONV_PROBE=3 python tools/conv_cases.py b64_c1x1_80,b64_c1x1_160,b64_c1x1_320 5 2>&1 | grep -v "^\["
