#!/bin/bash
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80
echo "== debug"; MDB_CONV_DEBUG=1 python tools/conv_cases.py $C 2 2>&1 | tail -16
for pr in 4 5 6 7 12 20; do echo "== probe $pr"; MDB_CONV_PROBE=$pr timeout 300 python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["; done
for pm in 256 64 0; do echo "== promo $pm"; MDB_TMAP_PROMO=$pm python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["; MDB_TMAP_PROMO=$pm MDB_CONV_PROBE=4 python tools/conv_cases.py $C 5 2>&1 | grep -v "^\[";  done
