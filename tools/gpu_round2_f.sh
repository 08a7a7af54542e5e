#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_aux.py tests/test_gpu_trace.py tests/test_gpu_e2e.py tests/test_gpu_equivalence.py -m gpu -q 2>&1 | tail -8) > gpurun_out/r02_gputest_f.log; tail -8 gpurun_out/r02_gputest_f.log | cut -c1-400
MDB_LIB_VARIANT=nmst MDB_NMS_DEBUG=1 timeout 300 python tools/profile_one_pass.py 64 2>&1 | grep "nms dbg\|timing" | head -4
timeout 300 python bench.py --no-cpu-baseline --no-incumbent 2> /dev/null > gpurun_out/r02_bench_f.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_f.json').read())
print('value %.0f e2e %.0f (%.2f ms) net %.2f post %.2f pre %.2f clocks %s e2edev %s'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['stage_ms_per_step']['network_K2_K4'],d['stage_ms_per_step']['decode_nms_K5_K8'],d['stage_ms_per_step']['preprocess_K1'],d['clocks']['sm_mhz'],d['e2e']['device_ms_per_step']))
PY
timeout 300 python tools/profile_layers.py 64 gpurun_out/r02_layers_bs64.csv | grep -i "upsample\|sppf\|TOTAL\|images/s"
