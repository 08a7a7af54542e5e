#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3
for split in 1 2 4; do
MDB_SPLIT=$split timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split$split.json 2> gpurun_out/bench.err; echo "bench split $split exit $?"
python - <<PY
import json
d = json.load(open('gpurun_out/bench_split$split.json'))
print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches']}, d['stage_ms_per_step'], {k: d['e2e'][k] for k in ['value','ms_per_step','device_ms_per_step']}, round(d['roofline']['frac'],3), d['clocks'])
PY
done
