#!/bin/bash
# Round-2 final evidence pass (GPU box): whole GPU test suite, smoke, the bench lines of configs 2/3/4 (+ cpu baseline + incumbent),
# launch list of the bench command, DRAM traffic of one resident step, full ncu captures of the kernels rewritten this round.
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r02_gputest_final.log; tail -6 gpurun_out/r02_gputest_final.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py 2> gpurun_out/r02_bench_bs64_final.err | tee gpurun_out/r02_bench_bs64_final.json | cut -c1-200
timeout 600 python bench.py --batch 8 --steps 20 --no-cpu-baseline 2> /dev/null | tee gpurun_out/r02_bench_bs8.json | cut -c1-160
timeout 600 python bench.py --model 'synthetic://mdv5b' --no-cpu-baseline --no-incumbent 2> /dev/null | tee gpurun_out/r02_bench_mdv5b.json | cut -c1-160
export MDB_NO_GRAPHS=1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-incumbent > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches_bench.csv
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'conv_pair|letterbox|nms_image|decode_kernel|upsample2x|sppf_' -c 400 --csv --log-file gpurun_out/traffic_step_bs64.csv python tools/profile_one_pass.py 64 > gpurun_out/traffic_run.log 2>&1
echo "traffic exit $?"; wc -l gpurun_out/traffic_step_bs64.csv
unset MDB_NO_GRAPHS
for k in upsample2x sppf_plane nms_image decode_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/ncu_final_$k python tools/profile_one_pass.py 64 > gpurun_out/ncu_final_$k.log 2>&1; echo "$k capture exit $?"
done
for c in c3x3_320 c3x3_160 s2_80_160 c1x1_160; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 1 -c 1 -f -o gpurun_out/ncu_final_$c python tools/conv_cases.py b64_$c 2 > gpurun_out/ncu_final_$c.log 2>&1; echo "conv capture $c exit $?"
done
timeout 300 python tools/profile_layers.py 64 gpurun_out/r02_layers_bs64.csv > gpurun_out/profile_layers.log 2>&1; echo "layers exit $?"; tail -2 gpurun_out/profile_layers.log
bash tools/gpu_sanitizer.sh
