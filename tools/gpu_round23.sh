#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_letterbox.py tests/test_gpu_e2e.py tests/test_gpu_equivalence.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3 | cut -c1-400
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_cur.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches']}, d['stage_ms_per_step'], {k: d['e2e'][k] for k in ['value','ms_per_step','device_ms_per_step']}, round(d['roofline']['frac'],3), d['clocks'])
PY
