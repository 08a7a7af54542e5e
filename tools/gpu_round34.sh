#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80,b64_c3x3_160,b64_c3x3_320,c1x1_640,s2_80_160,stem,c3x3_480
for i in 1 2; do
echo "== prev"; MDB_LIB_VARIANT=prev python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
echo "== new"; python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
done
