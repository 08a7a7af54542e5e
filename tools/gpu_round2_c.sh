#!/bin/bash
# Round-2 pass C: NMS v2.1 + JPEG path validation, polling back-off A/B
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_jpeg.py tests/test_gpu_e2e.py tests/test_gpu_equivalence.py -m gpu -q -s 2>&1 | grep -v "^Loading\|^No image\|^Using model\|^PTDetector" | tail -30) > gpurun_out/r02_gputest_c.log; tail -30 gpurun_out/r02_gputest_c.log | cut -c1-700
for bo in 0 40 100 250; do
  MDB_CONV_BACKOFF=$bo timeout 300 python bench.py --no-cpu-baseline --no-incumbent 2> /dev/null > gpurun_out/r02_bench_bo$bo.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_bo$bo.json').read())
print('backoff $bo: value %.0f e2e %.0f (%.2f ms) net %.2f post %.2f  clocks %s'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['stage_ms_per_step']['network_K2_K4'],d['stage_ms_per_step']['decode_nms_K5_K8'],d['clocks']['sm_mhz']))
PY
done
for bo in 0 100; do echo "backoff $bo"; MDB_CONV_BACKOFF=$bo timeout 200 python tools/conv_cases.py b64_c1x1_160,b64_c1x1_80,b64_c1x1_320,b64_c1x1_160_80 20; done
