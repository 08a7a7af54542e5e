#!/bin/bash
mkdir -p gpurun_out
echo "== NMS phase timers (variant build)"
MDB_LIB_VARIANT=nmst MDB_NMS_DEBUG=1 timeout 300 python tools/profile_one_pass.py 64 2>&1 | grep "nms dbg\|timing" | head -8
echo "== edge taper A/B"
for et in 0 1 0 1; do
  MDB_EDGE_TAPER=$et timeout 300 python bench.py --no-cpu-baseline --no-incumbent 2> /dev/null > gpurun_out/r02_bench_et$et.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_et$et.json').read())
print('taper $et: value %.0f e2e %.0f (%.2f ms) net %.2f post %.2f clocks %s e2edev %s'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['stage_ms_per_step']['network_K2_K4'],d['stage_ms_per_step']['decode_nms_K5_K8'],d['clocks']['sm_mhz'],d['e2e']['device_ms_per_step']))
PY
done
