#!/bin/bash
# Round-2 pass B: new NMS validation, bench, source-level ncu capture of a 1x1 layer (epilogue stall reasons)
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_nms.py tests/test_gpu_e2e.py tests/test_gpu_equivalence.py tests/test_gpu_aux.py tests/test_gpu_parity_1e3.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/r02_gputest_nms.log; tail -6 gpurun_out/r02_gputest_nms.log
timeout 600 python bench.py --no-cpu-baseline --no-incumbent 2> gpurun_out/r02_bench_nms.err | tee gpurun_out/r02_bench_nms.json | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_nms.json').read())
print('value %.0f e2e %.0f (%.2f ms) stage %s e2edev %s'%(d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['stage_ms_per_step'],d['e2e']['device_ms_per_step']))
PY
timeout 300 python tools/conv_cases.py b64_c1x1_160,b64_c1x1_80,b64_c1x1_320,b64_c3x3_320 20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 1 -c 1 -f -o gpurun_out/r02_ncu_c1x1_160 python tools/conv_cases.py b64_c1x1_160 2 > gpurun_out/r02_ncu_c1x1_160.log 2>&1; echo "1x1 capture exit $?"
