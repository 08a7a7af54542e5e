#!/bin/bash
# mid-point validation of the new defaults (tail-k, single-box halo, PDL, small-batch head chunk) + A/B of each in the bench
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6)
run() {
  env $1 timeout 300 python bench.py --no-cpu-baseline --no-incumbent $2 2> /dev/null > gpurun_out/r02_bench_p.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_p.json').read())
print('$1 $2: value %.0f e2e %.0f net %.2f clocks %s power %s frac %.3f'%(d['value'],d['e2e']['value'],d['stage_ms_per_step']['network_K2_K4'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max'),d['roofline']['frac']))
PY
}
run MDB_CONV_HALO1=0 ""; run MDB_CONV_HALO1=1 ""; run MDB_CONV_HALO1=0 ""; run MDB_CONV_HALO1=1 ""
run MDB_SMALL_SPLIT=0 "--batch 8"; run MDB_SMALL_SPLIT=1 "--batch 8"; run MDB_SMALL_SPLIT=0 "--batch 8"; run MDB_SMALL_SPLIT=1 "--batch 8"
