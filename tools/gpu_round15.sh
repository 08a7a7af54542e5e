#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_trace.py 2>&1 | grep -E "consume|generate" | tee gpurun_out/e2e_trace.log
timeout 300 python tools/conv_cases.py p2_c64,p2_c80,p2_c128,stem,stem_c64 2>&1 | tee gpurun_out/conv_oob.log
