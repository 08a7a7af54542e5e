#!/bin/bash
# same-box A/B of the whole network: old tree (_old = dfb3e1c~1) vs the current tree with single features switched off
mkdir -p gpurun_out
run() {
  (cd $1 && env $2 timeout 300 python bench.py --no-cpu-baseline --no-incumbent --steps 12 2> /dev/null > /tmp/b.json; python - <<PY
import json
d=json.loads(open('/tmp/b.json').read())
print('%-6s %-22s value %.0f e2e %.0f net %.2f clocks %s power %s'%('$1','$2',d['value'],d['e2e']['value'],d['stage_ms_per_step']['network_K2_K4'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max')))
PY
)
}
for round in 1 2 3; do
  run _old X=0
  run . X=0
  run . MDB_CONV_PDL=0
  run . MDB_CONV_HALO_S2=0
  run . MDB_CONV_HALO1=0
  run . MDB_CONV_TAILK=0
done
