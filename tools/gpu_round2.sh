#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --no-header -s -p no:cacheprovider > gpurun_out/test_gpu_e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/summary.txt; tail -n 15 gpurun_out/test_gpu_e2e.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -2 gpurun_out/smoke.log
timeout 600 python tools/profile_layers.py 64 gpurun_out/layers_bs64.csv > gpurun_out/profile.log 2>&1; echo "profile exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/profile.log
python - <<'PY' > gpurun_out/cpu_threads.log 2>&1
import time, torch, os
import torch.nn.functional as F
print('cpus', os.cpu_count())
x = torch.rand(1, 320, 96, 160); w = torch.randn(320, 320, 3, 3)
for nt in [8, 16, 32, 64, 128]:
    torch.set_num_threads(nt)
    F.conv2d(x, w, padding=1)
    t = time.time()
    for _ in range(3): F.conv2d(x, w, padding=1)
    dt = (time.time() - t) / 3
    print(nt, 'threads: %.1f ms  %.0f GFLOP/s' % (dt * 1e3, 2 * 96 * 160 * 320 * 320 * 9 / dt / 1e9))
PY
cat gpurun_out/cpu_threads.log
cat gpurun_out/summary.txt
