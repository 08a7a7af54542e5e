#!/bin/bash
# bisect of the in-network regression: old tree (644bda5), tail-k commit (dfb3e1c), current, current with the filter-row loop not unrolled
mkdir -p gpurun_out
run() {
  (cd $1 && env $2 timeout 300 python bench.py --no-cpu-baseline --no-incumbent 2> /dev/null > /tmp/b.json; python - <<PY
import json
d=json.loads(open('/tmp/b.json').read())
print('%-6s %-22s value %.0f e2e %.0f net %.2f clocks %s power %s'%('$1','$2',d['value'],d['e2e']['value'],d['stage_ms_per_step']['network_K2_K4'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max')))
PY
)
}
for round in 1 2 3; do
  run _old X=0
  run _c1 X=0
  run . X=0
  run . MDB_LIB_VARIANT=kh1
done
