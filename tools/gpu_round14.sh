#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_conv.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python tools/e2e_profile.py 2>&1 | grep -E "ms per call|infer_batch ms" 
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open('gpurun_out/bench_cur.json'))
print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches']}, d['stage_ms_per_step'], {k: d['e2e'][k] for k in ['value','ms_per_step','device_ms_per_step']}, round(d['roofline']['frac'],3), d['clocks'])
PY
