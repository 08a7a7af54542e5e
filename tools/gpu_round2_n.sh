#!/bin/bash
mkdir -p gpurun_out
CASES=b64_c1x1_160,b64_c1x1_80,b64_c1x1_320,b64_c1x1_160_80,b64_c3x3_320,b64_c3x3_160,b64_c3x3_80,stem
echo "== tail-k skip off / on (default lib)"
MDB_CONV_TAILK=0 timeout 200 python tools/conv_cases.py $CASES 20 2>&1 | grep -v "^Loading"
timeout 200 python tools/conv_cases.py $CASES 20 2>&1 | grep -v "^Loading"
for v in cw32w16 cw32w20; do
  echo "== variant '${v}'"
  MDB_LIB_VARIANT=$v timeout 200 python tools/conv_cases.py $CASES 20 2>&1 | grep -v "^Loading"
done
echo "== role timers 3x3 80 / 160 (tail-k on, then off)"
MDB_LIB_VARIANT=convt MDB_CONV_DEBUG=1 timeout 200 python tools/conv_cases.py b64_c3x3_80,b64_c3x3_160 3 2>&1 | grep -v "^Loading" | tail -12
MDB_CONV_TAILK=0 MDB_LIB_VARIANT=convt MDB_CONV_DEBUG=1 timeout 200 python tools/conv_cases.py b64_c3x3_80,b64_c3x3_160 3 2>&1 | grep -v "^Loading" | tail -12
echo "== correctness (default lib with tail-k, then cw32w20)"
(timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_trace.py -m gpu -q 2>&1 | tail -3)
(MDB_LIB_VARIANT=cw32w20 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_trace.py::test_every_layer_of_yolov5x6_against_pytorch_on_the_engines_own_inputs tests/test_gpu_trace.py::test_every_layer_of_the_tiny_model_mixed_batch -m gpu -q 2>&1 | tail -3)
run() {
  MDB_CONV_TAILK=$2 MDB_LIB_VARIANT=$1 timeout 300 python bench.py --no-cpu-baseline --no-incumbent 2> /dev/null > gpurun_out/r02_bench_n.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n.json').read())
print('variant "$1" tailk=$2: value %.0f e2e %.0f net %.2f clocks %s power %s frac %.3f'%(d['value'],d['e2e']['value'],d['stage_ms_per_step']['network_K2_K4'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max'),d['roofline']['frac']))
PY
}
run "" 0; run "" 1; run cw32w20 1; run cw32w16 1; run "" 0; run "" 1; run cw32w20 1; run cw32w16 1
