#!/bin/bash
mkdir -p gpurun_out
for v in "" cw8w24 cw8w16 cw8w28; do
  echo "== variant '${v}'"
  MDB_LIB_VARIANT=$v timeout 200 python tools/conv_cases.py b64_c1x1_160,b64_c1x1_80,b64_c1x1_320,b64_c1x1_160_80,b64_c3x3_320,b64_c3x3_80,stem 20 2>&1 | grep -v "^Loading"
done
echo "== correctness of cw8w24"
(MDB_LIB_VARIANT=cw8w24 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_trace.py::test_every_layer_of_yolov5x6_against_pytorch_on_the_engines_own_inputs tests/test_gpu_trace.py::test_every_layer_of_the_tiny_model_mixed_batch -m gpu -q 2>&1 | tail -3)
for v in "" cw8w24 "" cw8w24; do
  MDB_LIB_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline --no-incumbent 2> /dev/null > gpurun_out/r02_bench_m.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_m.json').read())
print('variant "$v": value %.0f e2e %.0f net %.2f clocks %s power %s frac %.3f'%(d['value'],d['e2e']['value'],d['stage_ms_per_step']['network_K2_K4'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max'),d['roofline']['frac']))
PY
done
