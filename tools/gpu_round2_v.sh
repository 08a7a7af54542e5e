#!/bin/bash
# per-layer same-box diff of single features
mkdir -p gpurun_out
(cd _old && timeout 300 python tools/profile_layers.py 64 ../gpurun_out/layers_v_old.csv > /dev/null 2>&1; echo "old exit $?")
for cfg in X=0 MDB_CONV_HALO1=0 MDB_CONV_TAILK=0 MDB_CONV_HALO_S2=0 MDB_CONV_PDL=0 "MDB_CONV_HALO1=0 MDB_CONV_TAILK=0 MDB_CONV_HALO_S2=0 MDB_CONV_PDL=0"; do
  name=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python tools/profile_layers.py 64 gpurun_out/layers_v_$name.csv > /dev/null 2>&1; echo "$name exit $?"
done
