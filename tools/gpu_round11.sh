#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_layers.py 64 gpurun_out/layers_bs64_cur.csv > gpurun_out/profile_cur.log 2>&1; tail -2 gpurun_out/profile_cur.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_cur.json'))
print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches', 'stage_ms_per_step']}, d['e2e'], d['roofline']['frac'], d['clocks'])
PY
