#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
C=b64_c1x1_80,b64_c1x1_160,b64_c1x1_320,b64_c3x3_80,b64_c3x3_160,b64_c3x3_320,c1x1_640,s2_80_160
for v in "" early; do echo "== variant '$v'"; MDB_LIB_VARIANT=$v python tools/conv_cases.py $C 5 2>&1 | grep -v "^\["; done
B='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["gpu_launches"], d["clocks"])'
for i in 1 2; do
echo "== bench nofuse"; MDB_FUSE_C3=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$B"
echo "== bench fuse"; timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench40_$i.json | python -c "$B"
echo "== bench fuse+early"; MDB_LIB_VARIANT=early timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$B"
done
