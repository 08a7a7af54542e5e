"""
Design experiment for the CONTRACTIVE synthetic weight set ("next round" item 1a).

Runs on the CPU: draws a weight set with the given calibration targets, then measures
  * the relative perturbation gain stem -> Detect inputs (inject a 1e-3 relative perturbation into the stem input),
  * the error of a PyTorch model of the engine's arithmetic contract (fp16 operands + fp16 storage, fp32 accumulate)
    against the fp32 oracle: scores and boxes,
  * how degenerate the predictions are (spread of the logits over anchors, candidates above the thresholds).

    python tools/experiments/contractive.py --arch tiny --pre 0.5 --res 0.3 --bias 0.5
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from megadetector_b200 import synth_weights as sw  # noqa: E402
from megadetector_b200.synth import synth_image  # noqa: E402
from oracle import letterbox as olb  # noqa: E402
from oracle import yolov5 as oy  # noqa: E402


def h(t):
    return t.half().float()


def contract_forward(x, w, arch):
    """oracle.yolov5.forward with the engine's contract patched in (see tests/test_gpu_e2e.py::_fp16_contract_model)."""
    def conv_model(x, w, name, s=1, p=0, act=True, trace=None):
        y = F.conv2d(h(x), h(w[name + '.weight']), w[name + '.bias'], stride=s, padding=p)
        if act:
            y = y * torch.sigmoid(y)
        y = h(y) if act else y
        if trace is not None:
            trace[name] = y
        return y

    def c3_model(x, w, i, n, shortcut, trace):
        pre = 'model.%d' % i
        a = conv_model(x, w, pre + '.cv1.conv')
        for j in range(n):
            b = conv_model(a, w, '%s.m.%d.cv1.conv' % (pre, j))
            nm = '%s.m.%d.cv2.conv' % (pre, j)
            y = F.conv2d(h(b), h(w[nm + '.weight']), w[nm + '.bias'], stride=1, padding=1)
            y = y * torch.sigmoid(y)
            a = h(a + y) if shortcut else h(y)
        c = conv_model(x, w, pre + '.cv2.conv')
        return conv_model(torch.cat((a, c), 1), w, pre + '.cv3.conv')

    o1, o2 = oy._conv, oy._c3
    oy._conv, oy._c3 = conv_model, c3_model
    try:
        return oy.forward(x, w, arch)
    finally:
        oy._conv, oy._c3 = o1, o2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--arch', default='tiny')
    ap.add_argument('--profile', default='contractive')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--images', type=int, default=2)
    ap.add_argument('--hw', default='1080x1920')
    args = ap.parse_args()
    torch.set_num_threads(8)
    a = sw.YOLOV5X6 if args.arch != 'tiny' else sw.TINY_P6
    arch = oy.Arch() if args.arch != 'tiny' else oy.tiny_arch()
    t0 = time.time()
    w = sw.synth_yolov5_weights(a['ch'], a['n_backbone'], a['n_head'], a['nc'], seed=args.seed, profile=args.profile)
    print('weights: %.1fs' % (time.time() - t0))
    H, W = (int(v) for v in args.hw.split('x'))
    for i in range(args.images):
        img = synth_image(i, H, W)
        lb, _, _ = olb.letterbox(img, 1280, 64, auto=True, scaleup=True)
        x = torch.from_numpy(olb.to_model_input(lb))[None]
        tr = {}
        t0 = time.time()
        p32 = oy.forward(x, w, arch, trace=tr)
        t_fwd = time.time() - t0
        # perturbation gain: relative 1e-3 white perturbation of the stem input
        g = torch.Generator().manual_seed(99)
        dx = torch.randn(x.shape, generator=g) * 1e-3 * x.std()
        tr2 = {}
        oy.forward(x + dx, w, arch, trace=tr2)
        rel_in = float(dx.norm() / (x - x.mean()).norm())
        gains = {}
        for name in ('model.0.conv', 'model.2.cv3.conv', 'model.4.cv3.conv', 'model.6.cv3.conv', 'model.10.cv3.conv',
                     'model.23.cv3.conv', 'model.26.cv3.conv', 'model.29.cv3.conv', 'model.32.cv3.conv'):
            f0, f1 = tr[name], tr2[name]
            var = f0 - f0.mean(dim=(0, 2, 3), keepdim=True)  # the image-dependent (spatially varying) part
            gains[name] = (float((f1 - f0).norm() / f0.norm()) / rel_in, float((f1 - f0).norm() / var.norm()) / rel_in,
                           float(var.norm() / f0.norm()))
        pc = contract_forward(x, w, arch)
        e = (pc - p32).abs()[0]
        p = p32[0]
        obj, cls = p[:, 4], p[:, 5:]
        conf = (cls * obj[:, None]).max(1).values
        lvl = np.cumsum([0] + [3 * (lb.shape[0] // s) * (lb.shape[1] // s) for s in (8, 16, 32, 64)])
        print('image %d  fwd %.1fs  stem rel perturbation %.2e' % (i, t_fwd, rel_in))
        print('  gain (rel. to total | rel. to varying part | varying fraction):')
        for k, v in gains.items():
            print('    %-20s %.3f | %.3f | %.3f' % (k, v[0], v[1], v[2]))
        print('  contract vs fp32: max score err %.2e (obj %.2e cls %.2e)  max box err px %.3f  per level box %s'
              % (float(e[:, 4:].max()), float(e[:, 4].max()), float(e[:, 5:].max()), float(e[:, :4].max()),
                 ['%.3f' % float(e[lvl[k]:lvl[k + 1], :4].max()) for k in range(4)]))
        print('  obj: max %.3g  >1e-5: %d  >0.005: %d  >0.1: %d | conf>1e-5: %d  conf>0.05: %d conf>0.2: %d'
              % (float(obj.max()), int((obj > 1e-5).sum()), int((obj > 0.005).sum()), int((obj > 0.1).sum()),
                 int((conf > 1e-5).sum()), int((conf > 0.05).sum()), int((conf > 0.2).sum())))
        lo = torch.log(obj.clamp_min(1e-30) / (1 - obj).clamp_min(1e-30))
        print('  obj logit mean %.2f std %.2f | cls score mean %.3f std %.3f | wh/anchor spread: %s'
              % (float(lo.mean()), float(lo.std()), float(cls.mean()), float(cls.std()),
                 ['%.2f' % float(v) for v in torch.quantile(p[:, 2], torch.tensor([0.01, 0.5, 0.99]))]))


if __name__ == '__main__':
    main()
