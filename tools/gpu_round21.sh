#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -8 | cut -c1-700
echo "== tails on"; timeout 300 python tools/conv_cases.py c1x1_80,c3x3_80,s2_80_160,c1x1_160,c3x3_160,c3x3_480 2>&1
echo "== tails off"; MDB_CONV_TAIL=0 timeout 300 python tools/conv_cases.py c1x1_80,c3x3_80,s2_80_160,c1x1_160,c3x3_160,c3x3_480 2>&1
