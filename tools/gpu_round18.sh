#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_letterbox.py tests/test_gpu_conv.py tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider -s 2>&1 | grep -E "e2e\]|passed|failed|Error|error" | cut -c1-600
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python tools/profile_layers.py 64 gpurun_out/layers_bs64_cur.csv > gpurun_out/profile_cur.log 2>&1; head -3 gpurun_out/profile_cur.log | cut -c1-200; tail -2 gpurun_out/profile_cur.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cur.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_cur.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches']}, d['stage_ms_per_step'], {k: d['e2e'][k] for k in ['value','ms_per_step','device_ms_per_step']}, round(d['roofline']['frac'],3), d['clocks'])
PY
