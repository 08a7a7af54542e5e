#!/bin/bash
mkdir -p gpurun_out
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80,b64_c3x3_160,b64_c3x3_320,c1x1_640,s2_80_160,stem,c3x3_480
for v in prev "" epiold late; do echo "== variant '$v'"; MDB_LIB_VARIANT=$v python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["; done
MDB_LIB_VARIANT=epiold timeout 300 python -m pytest tests/test_gpu_conv.py -q -x 2>&1 | tail -2
for v in "" epiold "" epiold; do echo "== bench '$v'"; MDB_LIB_VARIANT=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['clocks'])"; done
