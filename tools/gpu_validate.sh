#!/bin/bash
# Round-end sanity on the GPU box: GPU tests (verbose report lines), smoke, default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "passed|failed|error|e2e\]|Error|assert" | cut -c1-1500 | tail -12
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_validate.err | tee gpurun_out/bench_validate.json | cut -c1-200
