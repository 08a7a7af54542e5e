#!/bin/bash
# per-layer same-box diff: old tree vs current
mkdir -p gpurun_out
for r in 1 2; do
  (cd _old && timeout 300 python tools/profile_layers.py 64 ../gpurun_out/layers_old_$r.csv > /dev/null 2>&1; echo "old $r exit $?")
  timeout 300 python tools/profile_layers.py 64 gpurun_out/layers_new_$r.csv > /dev/null 2>&1; echo "new $r exit $?"
done
tail -1 gpurun_out/layers_old_1.csv gpurun_out/layers_new_1.csv gpurun_out/layers_old_2.csv gpurun_out/layers_new_2.csv
