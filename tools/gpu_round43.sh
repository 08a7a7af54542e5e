#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py -q -x 2>&1 | tail -2
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80,b64_c3x3_160,b64_c3x3_320,c1x1_640,s2_80_160,stem
for v in prev "" prev ""; do echo "== variant '$v'"; MDB_LIB_VARIANT=$v python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["; done
B='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["clocks"])'
for v in prev "" prev ""; do echo "== bench '$v'"; MDB_LIB_VARIANT=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$B"; done
