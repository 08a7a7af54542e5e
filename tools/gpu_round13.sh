#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/conv_cases.py c1x1_160,c1x1_160_noact,c1x1_80,c1x1_320,c1x1_640 2>&1 | tee gpurun_out/conv_1x1.log
timeout 600 python tools/e2e_profile.py 2>&1 | grep -v "^Loading\|^No image\|^Using\|^PTDetector" | tee gpurun_out/e2e_profile.log | cut -c1-180
