#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
B='import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"]["value"], d["stage_ms_per_step"], d["gpu_launches"], d["clocks"])'
for i in 1 2; do timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench41_$i.json | python -c "$B"; done
python tools/profile_layers.py 64 gpurun_out/layers_bs64_v4.csv | tail -2
