#!/bin/bash
mkdir -p gpurun_out
./tools/micro/silu_bench
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench32.err | tee gpurun_out/bench32.json | cut -c1-400
python tools/profile_layers.py 64 gpurun_out/layers_bs64_v3.csv | tail -3
