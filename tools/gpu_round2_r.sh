#!/bin/bash
# same-box A/B: the tree before this session's conv changes (_old, commit dfb3e1c~1) against the current one
mkdir -p gpurun_out
CASES=b64_c1x1_160,b64_c1x1_320,c1x1_640,b64_c3x3_320,b64_c3x3_160,b64_c3x3_80,c3x3_480,b64_s2_80_160,b64_s2_320_640,s2_640_640,c1x1_1280_640
for round in 1 2; do
  echo "== old tree"
  (cd _old && timeout 300 python ../tools/conv_cases_ab.py . $CASES 20 2>&1 | grep -v "^Loading")
  echo "== current tree"
  timeout 300 python tools/conv_cases_ab.py . $CASES 20 2>&1 | grep -v "^Loading"
done
run() {
  (cd $1 && timeout 300 python bench.py --no-cpu-baseline --no-incumbent 2> /dev/null > /tmp/b.json; python - <<PY
import json
d=json.loads(open('/tmp/b.json').read())
print('$1: value %.0f e2e %.0f net %.2f clocks %s power %s frac %.3f'%(d['value'],d['e2e']['value'],d['stage_ms_per_step']['network_K2_K4'],d['clocks']['sm_mhz'],d['clocks'].get('power_w_max'),d['roofline']['frac']))
PY
)
}
run _old; run .; run _old; run .
