#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -12 | cut -c1-600
echo "== halo on"; timeout 300 python tools/conv_cases.py c3x3_80,c3x3_160,c3x3_320,c3x3_480,stem 2>&1 | tee gpurun_out/conv_halo.log
echo "== halo off"; MDB_CONV_HALO=0 timeout 300 python tools/conv_cases.py c3x3_80,c3x3_160,c3x3_320,c3x3_480,stem 2>&1
