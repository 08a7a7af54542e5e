#!/bin/bash
# compute-sanitizer memcheck over the engine's kernels (K2 on two layer shapes; then K1..K8 through the tiny model)
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python tools/conv_cases.py c1x1_160,c3x3_160,s2_80_160 2 > gpurun_out/sanitizer_conv.log 2>&1; echo "conv memcheck exit $?"
tail -6 gpurun_out/sanitizer_conv.log
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke > gpurun_out/sanitizer_smoke.log 2>&1; echo "smoke memcheck exit $?"
tail -6 gpurun_out/sanitizer_smoke.log
