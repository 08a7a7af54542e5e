#!/bin/bash
# after hoisting the tail test out of the per-MMA code: correctness, then same-box A/B against the old tree
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_trace.py -m gpu -q 2>&1 | tail -3)
run() {
  (cd $1 && env $2 timeout 300 python bench.py --no-cpu-baseline --no-incumbent --steps 12 $3 2> /dev/null > /tmp/b.json; python - <<PY
import json
d=json.loads(open('/tmp/b.json').read())
print('%-6s %-22s %-10s value %.0f e2e %.0f net %.2f clocks %s power %s'%('$1','$2','$3',d['value'],d['e2e']['value'],d['stage_ms_per_step']['network_K2_K4'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max')))
PY
)
}
for round in 1 2 3; do
  run _old X=0
  run . X=0
  run . MDB_CONV_PDL=0
done
run _old X=0 "--batch 8"; run . X=0 "--batch 8"; run _old X=0 "--batch 8"; run . X=0 "--batch 8"
