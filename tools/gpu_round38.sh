#!/bin/bash
mkdir -p gpurun_out
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80,b64_c3x3_160,b64_c3x3_320,c1x1_640,s2_80_160,stem,c3x3_480
echo "== prev"; MDB_LIB_VARIANT=prev python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
echo "== new"; python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["
for i in 1 2; do
echo "== bench prev"; MDB_LIB_VARIANT=prev timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['clocks'])"
echo "== bench new"; timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench38_$i.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stage_ms_per_step'], d['clocks'])"
done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
