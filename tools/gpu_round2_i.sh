#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_trace.py tests/test_gpu_parity_1e3.py::test_yolov5x6_contractive_bs8 tests/test_video.py -m gpu -q 2>&1 | tail -6) > gpurun_out/r02_gputest_i.log; tail -6 gpurun_out/r02_gputest_i.log | cut -c1-300
timeout 200 python tools/conv_cases.py b64_c1x1_160,b64_c1x1_80,b64_c1x1_320,b64_c1x1_160_80,b64_c3x3_320,b64_c3x3_80 20
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-incumbent 2> /dev/null > gpurun_out/r02_bench_i$i.json
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_i$i.json').read())
print('value %.0f e2e %.0f (%.2f ms) net %.2f post %.2f pre %.2f clocks %s power %s frac %.3f'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['stage_ms_per_step']['network_K2_K4'],d['stage_ms_per_step']['decode_nms_K5_K8'],d['stage_ms_per_step']['preprocess_K1'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max'),d['roofline']['frac']))
PY
done
timeout 300 python tools/profile_layers.py 64 gpurun_out/r02_layers_bs64_f32x2.csv | tail -2
