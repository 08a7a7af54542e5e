#!/bin/bash
# Round-2 evidence pass A (GPU box): whole GPU test suite, smoke, bench lines of configs 2/3/4 (+ incumbent), ncu captures of the
# non-conv kernels at the bench configuration.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r02_gputest_full.log; tail -5 gpurun_out/r02_gputest_full.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/r02_bench_bs64.err | tee gpurun_out/r02_bench_bs64.json | cut -c1-400
timeout 600 python bench.py --batch 8 --steps 20 --no-cpu-baseline 2> gpurun_out/r02_bench_bs8.err | tee gpurun_out/r02_bench_bs8.json | cut -c1-300
timeout 600 python bench.py --model 'synthetic://mdv5b' --no-cpu-baseline --no-incumbent 2> gpurun_out/r02_bench_mdv5b.err | tee gpurun_out/r02_bench_mdv5b.json | cut -c1-300
for k in letterbox decode_kernel upsample2x sppf_pool nms_image; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02_ncu_$k python tools/profile_one_pass.py 64 > gpurun_out/r02_ncu_$k.log 2>&1; echo "$k capture exit $?"
done
ls -la gpurun_out | grep r02 | head -40
