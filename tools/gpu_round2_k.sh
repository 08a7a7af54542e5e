#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_equivalence.py tests/test_gpu_letterbox.py tests/test_gpu_e2e.py::test_failure_isolation_and_dict_inputs tests/test_gpu_jpeg.py -m gpu -q 2>&1 | tail -4) > gpurun_out/r02_gputest_k.log; tail -4 gpurun_out/r02_gputest_k.log | cut -c1-300
timeout 300 python tools/e2e_pageable.py 2>/dev/null | grep images
timeout 900 python tools/bench_files_video.py 2> gpurun_out/r02_files_video.err | tee gpurun_out/r02_files_video.jsonl | cut -c1-330
