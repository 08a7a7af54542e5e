#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu exit $?"
tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
python - <<PY
import json
d = json.load(open('gpurun_out/bench_2gpu.json'))
print({k: d[k] for k in ['value', 'n_gpus', 'ms_per_step', 'gpu_launches']}, {k: d['e2e'][k] for k in ['value','ms_per_step']}, round(d['roofline']['frac'],3), d['clocks'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --ref-images 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cat gpurun_out/bench_ref.json | cut -c1-600
