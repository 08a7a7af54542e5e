#!/usr/bin/env python
"""
bench.py -- images/sec of the MDv5a hot path (BASELINE.json: "MDv5a 1280px bs=64"), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path (K1 letterbox .. K8 rescale + record formatting) over one batch of 64 synthetic
1920x1080 images per GPU (weak scaling: images are sharded over ranks as independent batches, no collective on
the data path -- SURVEY.md par. 8e).  Weights: calibrated random-init YOLOv5x6 (no MDv5a checkpoint offline).

JSON keys (see the task contract): `value` = whole-job device-resident throughput (inputs already in HBM; CUDA
events on the engine's compute stream; max over ranks), `e2e` = same metric through the reference-facing API
(`B200Detector.generate_detections_one_batch`) with pinned HOST images (H2D + result D2H + record formatting in
the timed region), `roofline` = conv stack (tcgen05 kernels) FLOP/s against the measured bf16 peak,
`cpu_baseline` = the CPU oracle of the same path on a bounded sample, `clocks` = nvidia-smi samples during the run.

`--impl reference` times the reference's CPU implementation of the path.  The reference is pure Python whose
hot-path arithmetic lives in absent third-party packages (ultralytics-yolov5), so it cannot be installed or run
here; the arm runs the CPU oracle port (oracle/), the one place besides `cpu_baseline` where bench.py executes it.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL = 'synthetic://yolov5x6?seed=0'
SRC_H, SRC_W = 1080, 1920
LB_H, LB_W = 768, 1280
CONF_THRES = 1e-5  # the batch path's detection_threshold (run_detector_batch.py:751-754)
GFLOP_PER_IMAGE = 498.97  # 2 x 249.485 GMAC at 768x1280 (SURVEY.md par. 8d / BASELINE.md par. 2)
POOL = 64  # distinct synthetic images per rank


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'tflops': float(d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1590.0))), 'hbm_gbs': float(d.get('hbm_gbs', 6650.0)),
                'source': 'measured (MEASURED_PEAKS.json, bf16_tflops_sustained)'}
    return {'tflops': 1400.0, 'hbm_gbs': 6650.0, 'source': 'fallback (B200_PROFILING.md)'}


def _conv_dram_traffic():
    """DRAM bytes (read + write) of the conv kernels over one forward pass of 64 images, from the committed ncu pass
    (profiles/*_dram_traffic_step_bs64.csv: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`)."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_dram_traffic_step_bs64.csv'))):
        for ln in open(path):
            if ln.startswith('conv_pair_kernel,'):
                _, n, rd, wr = ln.strip().split(',')
                best = (float(rd) + float(wr)) * 1e9, os.path.basename(path), int(n)  # the file holds exactly one pass
    return best


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
            'clocks_event_reasons.sw_power_cap'
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(',')]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for k, n in enumerate(names) if any(s[3 + k].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons,
                'power_w_max': max(float(s[2]) for s in self.samples), 'samples': len(sm)}


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's path on the host cores, bounded sample."""
    if rank != 0:
        return
    import torch
    from megadetector_b200 import checkpoint
    from megadetector_b200.synth import synth_image
    from oracle import yolov5 as oy
    from oracle.detector import OracleDetector
    cores = _best_cpu_threads()
    weights, _ = checkpoint.load_model(args.model)
    ora = OracleDetector(weights, oy.Arch(), fast_nms=True)
    per_step = max(1, args.ref_images)
    imgs = [synth_image(i, SRC_H, SRC_W) for i in range(per_step)]
    ids = ['img%05d.jpg' % i for i in range(per_step)]
    for _ in range(min(args.warmup, 1)):
        ora.generate_detections_one_batch(imgs[:1], ids[:1], CONF_THRES)
    t0 = time.time()
    n = 0
    for _ in range(args.steps):
        for i in range(per_step):  # the reference forces batch size 1 on CPU (run_detector_batch.py:1219-1223)
            ora.generate_detections_one_batch([imgs[i]], [ids[i]], CONF_THRES)
            n += 1
    dt = time.time() - t0
    v = n / dt
    line = {
        'impl': 'reference', 'metric': 'images/sec (MDv5a 1280px bs=64)', 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * dt / max(1, args.steps), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # the same workload as the engine's arm; each step is a bounded sample of it (the reference forces bs=1 on a CPU)
        'config': {'workload': 'MDv5a-shape YOLOv5x6 (calibrated random init, no checkpoint offline), bs={} per GPU, '
                               '1920x1080 RGB u8 -> letterbox 768x1280, conf_thres 1e-5, image-sharded'.format(args.batch),
                   'sample_images_per_step': per_step, 'reference_batch_size': 1},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                         'sample': '{} steps x {} synthetic 1920x1080 images, full path (letterbox, fp32 forward, nms, formatting), '
                                   'torch {} threads'.format(args.steps, per_step, torch.get_num_threads())},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=6)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--ref-images', type=int, default=4, help='images per step of the CPU reference arm')
    ap.add_argument('--cpu-baseline-images', type=int, default=6)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-incumbent', action='store_true', help='skip the stock PyTorch/cuDNN timing on the same GPU')
    ap.add_argument('--model', default=MODEL, help='model spec (config 4: synthetic://mdv5b) or a .pt file')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from megadetector_b200 import _capi
    from megadetector_b200.detector import B200Detector
    from megadetector_b200.synth import synth_image

    assert torch.cuda.is_available(), 'bench.py needs a B200; there is no CPU fallback'
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    W = max(3, args.warmup)
    det = B200Detector(args.model, {'max_batch': B, 'device': 'cuda:%d' % local_rank})
    eng = det.engine
    if os.environ.get('MDB_SPLIT'):
        eng.set_pipeline_split(int(os.environ['MDB_SPLIT']))
    # synthetic pool in page-locked host memory (each rank its own images: image-sharded job)
    pins = [_capi.PinnedArray((SRC_H, SRC_W, 3)) for _ in range(POOL)]
    for i, p in enumerate(pins):
        synth_image(rank * POOL + i, SRC_H, SRC_W, out=p.array)
    pool = [p.array for p in pins]
    ids = ['img%05d.jpg' % (rank * POOL + i) for i in range(POOL)]

    def batch(step):
        idx = [(step * B + j) % POOL for j in range(B)]
        return [pool[i] for i in idx], [ids[i] for i in idx]

    # ---------------- e2e: reference-facing API, host buffers in, records out ----------------
    for s in range(W):
        imgs, names = batch(s)
        res = det.generate_detections_one_batch(imgs, names, detection_threshold=CONF_THRES)
    n_dets = sum(len(r['detections']) for r in res)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count()
    barrier()
    t0 = time.perf_counter()
    dev_ms = {'preprocess_ms': 0.0, 'network_ms': 0.0, 'postprocess_ms': 0.0}
    for s in range(args.steps):
        imgs, names = batch(W + s)
        res = det.generate_detections_one_batch(imgs, names, detection_threshold=CONF_THRES)
        t = eng.last_timing()
        for k in dev_ms:
            dev_ms[k] += t[k]
    barrier()
    e2e_s = time.perf_counter() - t0
    launches_e2e = eng.launch_count() - launches0

    # ---------------- value: device-resident (the staged batch is replayed; CUDA events) ----------------
    eng.set_pipeline_split(1)
    imgs, names = batch(0)
    eng.infer_batch(imgs, CONF_THRES, want_anchors=False)  # stage one full batch in HBM
    eng.bench_resident(W, CONF_THRES, 7)
    barrier()
    launches1 = eng.launch_count()
    ms_all = eng.bench_resident(args.steps, CONF_THRES, 7)
    barrier()
    launches_res = eng.launch_count() - launches1
    ms_net = eng.bench_resident(args.steps, CONF_THRES, 2)
    ms_pre = eng.bench_resident(args.steps, CONF_THRES, 1)
    ms_post = eng.bench_resident(args.steps, CONF_THRES, 4)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t = torch.tensor([ms_all, e2e_s * 1000.0, ms_net, ms_pre, ms_post], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms, ms_net, ms_pre, ms_post = [float(x) for x in t.tolist()]

    if rank == 0:
        peaks = _peaks()
        total_imgs = B * args.steps * world
        value = total_imgs / (ms_all / 1000.0)
        e2e_v = total_imgs / (e2e_ms / 1000.0)
        conv_tflops = (B * args.steps * GFLOP_PER_IMAGE / 1000.0) / (ms_net / 1000.0)  # per GPU
        h2d = B * SRC_H * SRC_W * 3
        d2h = B * (300 * 6 * 4 + 4 + 300 * 4)
        line = {
            'metric': 'images/sec (MDv5a 1280px bs=64)', 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': W, 'ms_per_step': ms_all / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16 operands, f32 accumulate (tcgen05 kind::f16); u8 integer letterbox; f32 decode/NMS',
            'data': 'synthetic',
            'config': {'workload': 'MDv5a-shape YOLOv5x6 (calibrated random init, no checkpoint offline), bs={} per GPU, '
                                   '1920x1080 RGB u8 -> letterbox 768x1280, conf_thres 1e-5, image-sharded'.format(B),
                       'model': args.model, 'pool_images_per_rank': POOL, 'l2': 'inputs per step ({:.0f} MB u8 + activations) exceed the 126 MB L2'.format(h2d / 1e6),
                       'detections_last_batch': n_dets},
            'e2e': {'value': e2e_v, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_ms / args.steps, 'api': 'B200Detector.generate_detections_one_batch (pinned host images in, records out)',
                    'device_ms_per_step': {k: v / args.steps for k, v in dev_ms.items()}},
            'gpu_launches': int(launches_res),
            'gpu_launches_e2e': int(launches_e2e),
            'roofline': {'bound': 'tensor', 'achieved': conv_tflops, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                         'frac': conv_tflops / peaks['tflops'],
                         'traffic': (_conv_dram_traffic() or [None])[0] if B == 64 else None,
                         'traffic_source': (_conv_dram_traffic() or [None, None])[1], 'peak_source': peaks['source'],
                         'kernel': 'conv_pair_kernel<64> ({} launches/step; tcgen05.mma.cta_group::2) + 3 upsample + 1 SPPF pool launches'.format(int(launches_res) // max(1, args.steps) - 7),
                         'flops_per_step': B * GFLOP_PER_IMAGE * 1e9, 'ms_per_step': ms_net / args.steps},
            'stage_ms_per_step': {'preprocess_K1': ms_pre / args.steps, 'network_K2_K4': ms_net / args.steps,
                                  'decode_nms_K5_K8': ms_post / args.steps},
            'clocks': sampler.summary(),
        }
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(args.cpu_baseline_images, args.model)
    det.close()
    if rank == 0:
        if not args.no_incumbent and world == 1:
            # outside every timed region, after the engine has been destroyed: the incumbent on the same box
            try:
                line['incumbent'] = incumbent(args.model, B)
            except Exception as e:  # it is a reported side figure: never lose the bench line over it
                line['incumbent'] = {'error': str(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _best_cpu_threads():
    """torch's conv throughput on this box collapses when all hardware threads are used (measured on the 128-thread
    GPU host: 32 threads 1443 GFLOP/s, 128 threads 104 GFLOP/s), so the CPU arm uses the best of a few settings."""
    import torch
    import torch.nn.functional as F
    cores = os.cpu_count() or 1
    x = torch.rand(1, 320, 48, 80)
    w = torch.randn(320, 320, 3, 3)
    best, best_t = cores, None
    for nt in sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, 16), min(cores, 8)}, reverse=True):
        torch.set_num_threads(nt)
        F.conv2d(x, w, padding=1)
        t0 = time.time()
        for _ in range(3):
            F.conv2d(x, w, padding=1)
        dt = time.time() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
    torch.set_num_threads(best)
    return best


def incumbent(model, batch, iters=3):
    """
    BASELINE.md par. 4/6: what the reference's GPU path costs on THIS box -- the same module graph (oracle.yolov5.forward)
    through stock PyTorch/cuDNN on cuda:0, with torch's default TF32 convolutions (the reference's setting) and with
    `.half()` weights/activations.  Forward pass only (no letterbox, no NMS), inputs resident, CUDA events.  A reported
    side figure, timed after (and outside) every region of the engine's own numbers.
    """
    import torch
    from megadetector_b200 import checkpoint
    from oracle import yolov5 as oy
    weights, arch = checkpoint.load_model(model)
    a = oy.Arch(ch=list(arch['ch']), n_backbone=list(arch['n_backbone']), n_head=arch['n_head'], nc=arch['nc'])
    dev = torch.device('cuda')
    out = {'what': 'oracle.yolov5.forward (the reference\'s module graph) via stock PyTorch/cuDNN, forward only, '
                   'bs={} x 3x768x1280, {} timed iterations after 2 warm-ups'.format(batch, iters), 'torch': torch.__version__}
    for tag, dtype, tf32 in (('tf32', torch.float32, True), ('fp16', torch.float16, False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        wd = {k: v.to(dev, dtype) for k, v in weights.items()}
        x = torch.rand(batch, 3, LB_H, LB_W, device=dev, dtype=dtype)
        with torch.no_grad():
            for _ in range(2):
                oy.forward(x, wd, a)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                oy.forward(x, wd, a)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[tag] = {'ms_per_batch': ms, 'images_per_s': batch / (ms / 1000.0),
                    'conv_tflops': batch * GFLOP_PER_IMAGE / 1000.0 / (ms / 1000.0)}
        del wd, x
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return out


def cpu_baseline(n_images, model=MODEL):
    """The CPU oracle of the same path on the box's host cores, bounded sample (reported baseline, not a target)."""
    import torch
    from megadetector_b200 import checkpoint
    from megadetector_b200.synth import synth_image
    from oracle import yolov5 as oy
    from oracle.detector import OracleDetector
    cores = _best_cpu_threads()
    weights, _ = checkpoint.load_model(model)
    ora = OracleDetector(weights, oy.Arch(), fast_nms=True)
    imgs = [synth_image(i, SRC_H, SRC_W) for i in range(n_images + 1)]
    ora.generate_detections_one_batch(imgs[:1], ['w'], CONF_THRES)
    t0 = time.time()
    for i in range(1, n_images + 1):
        ora.generate_detections_one_batch([imgs[i]], ['i%d' % i], CONF_THRES)
    dt = time.time() - t0
    return {'value': n_images / dt, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
            'sample': '{} synthetic 1920x1080 images, bs=1, full path (oracle letterbox + fp32 torch forward + numpy nms + formatting), '
                      'torch {} threads (best of several settings on {} hardware threads), 1 warm-up'.format(n_images, torch.get_num_threads(), os.cpu_count())}


def _main_quiet():
    """
    bench.py's stdout must carry exactly one JSON line.  The drop-in detector prints the reference's load-time
    messages ("PTDetector using device ...", which AddaxAI parses) and NCCL prints its version banner from C, so
    file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor.
    """
    import contextlib
    import io
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    lines = []

    class _Tee(io.TextIOBase):
        def write(self, t):
            for ln in t.splitlines(True):
                if (ln.startswith('{"metric"') or ln.startswith('{"impl"')):
                    lines.append(ln)
                else:
                    sys.stderr.write(ln)
            return len(t)

    try:
        with contextlib.redirect_stdout(_Tee()):
            rc = main()
    finally:
        sys.stderr.flush()
        os.dup2(saved_fd, 1)
    for ln in lines:
        os.write(saved_fd, (ln if ln.endswith('\n') else ln + '\n').encode())
    os.close(saved_fd)
    return rc


if __name__ == '__main__':
    sys.exit(_main_quiet())
