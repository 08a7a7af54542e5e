#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_modern.py tests/test_gpu_jpeg.py tests/test_gpu_nms.py tests/test_gpu_letterbox.py tests/test_gpu_equivalence.py tests/test_video.py tests/test_gpu_e2e.py -m gpu -q -s 2>&1 | grep -v "^Loading\|^No image\|^Using model\|^PTDetector" | tail -60) > gpurun_out/r02_gputest_d.log; tail -60 gpurun_out/r02_gputest_d.log | cut -c1-900
timeout 300 python bench.py --no-cpu-baseline --no-incumbent 2> /dev/null > gpurun_out/r02_bench_d.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_d.json').read())
print('value %.0f e2e %.0f (%.2f ms) net %.2f post %.2f pre %.2f clocks %s e2edev %s'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['stage_ms_per_step']['network_K2_K4'],d['stage_ms_per_step']['decode_nms_K5_K8'],d['stage_ms_per_step']['preprocess_K1'],d['clocks']['sm_mhz'],d['e2e']['device_ms_per_step']))
PY
