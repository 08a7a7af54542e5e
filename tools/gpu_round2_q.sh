#!/bin/bash
# stride-2 halo main loop: correctness, then timing against the per-tap loop
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q 2>&1 | tail -4)
CASES=b64_s2_80_160,b64_s2_160_320,b64_s2_320_640,b64_s2_320_320,b64_c3x3_160
for h in 0 1; do
  echo "== MDB_CONV_HALO_S2=$h"
  MDB_CONV_HALO_S2=$h timeout 200 python tools/conv_cases.py $CASES 20 2>&1 | grep -v "^Loading"
done
(timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_parity_1e3.py -m gpu -q 2>&1 | tail -3)
run() {
  env $1 timeout 300 python bench.py --no-cpu-baseline --no-incumbent $2 2> /dev/null > gpurun_out/r02_bench_q.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_q.json').read())
print('$1 $2: value %.0f e2e %.0f net %.2f clocks %s power %s frac %.3f'%(d['value'],d['e2e']['value'],d['stage_ms_per_step']['network_K2_K4'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max'),d['roofline']['frac']))
PY
}
run MDB_CONV_HALO_S2=0 ""; run MDB_CONV_HALO_S2=1 ""; run MDB_CONV_HALO_S2=0 ""; run MDB_CONV_HALO_S2=1 ""
