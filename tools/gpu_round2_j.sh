#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_video.py tests/test_gpu_jpeg.py -m gpu -q 2>&1 | tail -5) > gpurun_out/r02_gputest_j.log; tail -5 gpurun_out/r02_gputest_j.log | cut -c1-300
timeout 300 python tools/e2e_pageable.py 2>/dev/null | grep images
timeout 900 python tools/bench_files_video.py 2> gpurun_out/r02_files_video.err | tee gpurun_out/r02_files_video.jsonl | cut -c1-330
