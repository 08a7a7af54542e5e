#!/bin/bash
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80,b64_c3x3_160,b64_c3x3_320,c1x1_640,s2_80_160,stem,c3x3_480
for i in 1 2; do
echo "== prev"; MDB_LIB_VARIANT=prev python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
echo "== new"; python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
done
echo "== ew20"; MDB_LIB_VARIANT=ew20 python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
echo "== timers"; MDB_LIB_VARIANT=timers MDB_CONV_DEBUG=1 python tools/conv_cases.py b64_c1x1_160,b64_c1x1_80,b64_c3x3_320 2 2>&1 | tail -9
