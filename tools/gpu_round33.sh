#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_conv.py -q -x 2>&1 | tail -3
C=b64_c3x3_80,c3x3_160,c3x3_320,c3x3_480,b64_c3x3_160,b64_c3x3_320
for i in 1 2; do
echo "== prev (1 tap/stage)"; MDB_LIB_VARIANT=prev python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
echo "== new (3 taps/stage)"; python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
done
echo "== new, halo for 80ch too"; MDB_CONV_HALO_MIN=2 python tools/conv_cases.py b64_c3x3_80,p2_c80,p2_c64,p2_c128 5 2>&1 | grep -v "^\["
echo "== new, no halo"; MDB_CONV_HALO=0 python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
