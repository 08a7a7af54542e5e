#!/bin/bash
# Final evidence pass (GPU box): launch list of the bench command, DRAM traffic of one resident step, full ncu capture of
# the dominant conv shape.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-300
export MDB_NO_GRAPHS=1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_bench.csv
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'conv_pair|letterbox|nms_image|decode_kernel|upsample2x|sppf_pool' -c 300 --csv --log-file gpurun_out/traffic_step_bs64.csv python tools/profile_layers.py 64 > gpurun_out/traffic_run.log 2>&1
echo "traffic exit $?"; wc -l gpurun_out/traffic_step_bs64.csv
unset MDB_NO_GRAPHS
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 1 -c 1 -f -o gpurun_out/ncu_final_c3x3_320 python tools/conv_cases.py c3x3_320 2 > gpurun_out/ncu_final.log 2>&1; echo "full capture exit $?"
timeout 600 ncu --set full --clock-control none -k regex:letterbox -c 1 -f -o gpurun_out/ncu_final_k1 python tools/e2e_profile.py > /dev/null 2>&1; echo "k1 capture exit $?"
timeout 600 ncu --set full --clock-control none -k regex:nms_image -c 1 -f -o gpurun_out/ncu_final_nms python tools/e2e_profile.py > /dev/null 2>&1; echo "nms capture exit $?"
ls -la gpurun_out | head -30
