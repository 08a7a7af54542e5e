#!/bin/bash
# last tree: conv / trace / parity tests, smoke, the bench lines (bs64 with cpu baseline + incumbent, bs8)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_trace.py tests/test_gpu_parity_1e3.py -m gpu -q 2>&1 | tail -3)
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py 2> gpurun_out/r02_bench_bs64_final.err | tee gpurun_out/r02_bench_bs64_final.json | cut -c1-200
timeout 600 python bench.py --batch 8 --steps 20 --no-cpu-baseline 2> /dev/null | tee gpurun_out/r02_bench_bs8.json | cut -c1-160
