#!/bin/bash
# single-box halo: does a 128-byte start offset / 1280-byte SBO address correctly?  (base-offset field 0, then filled)
mkdir -p gpurun_out
for h1 in 1 2; do
  echo "== MDB_CONV_HALO1=$h1 correctness"
  (MDB_CONV_HALO1=$h1 timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -4)
done
CASES=b64_c3x3_320,b64_c3x3_160,b64_c3x3_80,c3x3_480
for h1 in 0 1 2; do
  echo "== timing MDB_CONV_HALO1=$h1"
  MDB_CONV_HALO1=$h1 timeout 200 python tools/conv_cases.py $CASES 20 2>&1 | grep -v "^Loading"
done
echo "== HALO1=1 a_stages 2 / 4"
MDB_CONV_HALO1=1 MDB_CONV_HALO_ASTAGES=2 timeout 200 python tools/conv_cases.py $CASES 20 2>&1 | grep -v "^Loading"
MDB_CONV_HALO1=1 MDB_CONV_HALO_ASTAGES=4 timeout 200 python tools/conv_cases.py $CASES 20 2>&1 | grep -v "^Loading"
echo "== PDL correctness (conv + trace + parity bs8)"
(MDB_CONV_PDL=1 timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_trace.py "tests/test_gpu_parity_1e3.py" -m gpu -q -x 2>&1 | tail -4)
run() {
  env $1 timeout 300 python bench.py --no-cpu-baseline --no-incumbent $2 2> /dev/null > gpurun_out/r02_bench_o.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_o.json').read())
print('$1 $2: value %.0f e2e %.0f net %.2f clocks %s power %s frac %.3f'%(d['value'],d['e2e']['value'],d['stage_ms_per_step']['network_K2_K4'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max'),d['roofline']['frac']))
PY
}
run MDB_CONV_PDL=0 ""; run MDB_CONV_PDL=1 ""; run MDB_CONV_PDL=0 ""; run MDB_CONV_PDL=1 ""
run MDB_CONV_PDL=0 "--batch 8"; run MDB_CONV_PDL=1 "--batch 8"
