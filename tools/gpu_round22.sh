#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -8 | cut -c1-600
