#!/bin/bash
# st_post validation + corrected launch list / traffic passes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench24.err | tee gpurun_out/bench24.json
export MDB_NO_GRAPHS=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_bench.csv
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'conv_pair|letterbox|nms_image|decode_kernel|upsample2x|sppf_pool' -c 171 --csv --log-file gpurun_out/traffic_step_bs64.csv python tools/profile_layers.py 64 > gpurun_out/traffic_run.log 2>&1
echo "traffic exit $?"; wc -l gpurun_out/traffic_step_bs64.csv
