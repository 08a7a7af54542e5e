#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_e2e.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/conv_cases.py c1x1_80,c3x3_80,stem_c64,c1x1_160,c1x1_320,c3x3_320 2>&1
timeout 600 python tools/profile_layers.py 64 gpurun_out/layers_bs64_cur.csv > gpurun_out/profile_cur.log 2>&1; tail -2 gpurun_out/profile_cur.log
