#!/bin/bash
mkdir -p gpurun_out
MDB_CONV_DEBUG=1 timeout 300 python tools/conv_cases.py c1x1_160,c3x3_80,c3x3_160,c3x3_320,c3x3_480,stem 2 > gpurun_out/conv_dbg.log 2>&1; cat gpurun_out/conv_dbg.log
