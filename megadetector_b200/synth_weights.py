"""
Synthetic ("calibrated random-init") YOLOv5-P6 weights for benchmarks and tests.

BASELINE.json asks for MDv5a weights "from the reference mount (random-init YOLOv5x6 if unavailable)"; there is
no checkpoint offline, so this module manufactures one.  Plain random weights are useless for a ~100-conv deep
residual network without normalisation layers (BN is folded): activations explode or die, every anchor gets the
same score and threshold/NMS degenerate into pure tie-breaking (SURVEY.md par. 8d).  So the generator draws
N(0, 1/fan_in) weights and then walks the graph once over a letterboxed synthetic benchmark image, rescaling
each conv so that its pre-activation std is 1 (0.7 on residual branches; with SiLU that puts the stack close to the
edge of chaos, perturbation gain ~1.07 per layer) and shaping the Detect logits
(objectness ~ N(-13, 2.5): 10^3..10^4 anchors above the 1e-5 batch-path threshold, 10^1..10^2 above 0.005).

This is a workload generator (it plays the role of a checkpoint file), not part of the detection path and not
the parity oracle; it uses torch only to run the one calibration pass.  Output: dict name -> float32 torch
tensor in the fused-yolov5 state-dict naming (`model.{i}...conv.weight/bias`, `model.33.m.{l}.weight/bias`).
"""

import numpy as np
import torch
import torch.nn.functional as F

YOLOV5X6 = dict(ch=(80, 160, 320, 640, 960, 1280), n_backbone=(4, 8, 12, 4, 4), n_head=4, nc=3)
TINY_P6 = dict(ch=(32, 64, 128, 256, 384, 512), n_backbone=(1, 2, 2, 1, 1), n_head=1, nc=3)


def _specs(ch, n_backbone, n_head, nc):
    """(name, c_in, c_out, k) for every fused conv, execution order."""
    c = ch
    out = []

    def c3(i, ci, co, n):
        h = co // 2
        out.append(('model.%d.cv1.conv' % i, ci, h, 1))
        out.append(('model.%d.cv2.conv' % i, ci, h, 1))
        for j in range(n):
            out.append(('model.%d.m.%d.cv1.conv' % (i, j), h, h, 1))
            out.append(('model.%d.m.%d.cv2.conv' % (i, j), h, h, 3))
        out.append(('model.%d.cv3.conv' % i, 2 * h, co, 1))

    out.append(('model.0.conv', 3, c[0], 6))
    out.append(('model.1.conv', c[0], c[1], 3))
    c3(2, c[1], c[1], n_backbone[0])
    out.append(('model.3.conv', c[1], c[2], 3))
    c3(4, c[2], c[2], n_backbone[1])
    out.append(('model.5.conv', c[2], c[3], 3))
    c3(6, c[3], c[3], n_backbone[2])
    out.append(('model.7.conv', c[3], c[4], 3))
    c3(8, c[4], c[4], n_backbone[3])
    out.append(('model.9.conv', c[4], c[5], 3))
    c3(10, c[5], c[5], n_backbone[4])
    out.append(('model.11.cv1.conv', c[5], c[5] // 2, 1))
    out.append(('model.11.cv2.conv', c[5] * 2, c[5], 1))
    out.append(('model.12.conv', c[5], c[4], 1))
    c3(15, 2 * c[4], c[4], n_head)
    out.append(('model.16.conv', c[4], c[3], 1))
    c3(19, 2 * c[3], c[3], n_head)
    out.append(('model.20.conv', c[3], c[2], 1))
    c3(23, 2 * c[2], c[2], n_head)
    out.append(('model.24.conv', c[2], c[2], 3))
    c3(26, 2 * c[2], c[3], n_head)
    out.append(('model.27.conv', c[3], c[3], 3))
    c3(29, 2 * c[3], c[4], n_head)
    out.append(('model.30.conv', c[4], c[4], 3))
    c3(32, 2 * c[4], c[5], n_head)
    for l, ci in enumerate([c[2], c[3], c[4], c[5]]):
        out.append(('model.33.m.%d' % l, ci, 3 * (nc + 5), 1))
    return out


# Calibration targets per profile.  "chaotic" is the round-1 set (kept bit-for-bit: fixtures were produced with it);
# "contractive" is the set on which the 1e-3 parity bar is asserted (see the module docstring).
PROFILES = {
    'chaotic': dict(pre=1.0, res=0.7, bias=None, box_sd=1.0, cls=(0.0, 1.5), obj=(-13.0, 2.5)),
    'contractive': dict(pre=0.05, res=0.02, bias=0.0, box_sd=0.03, cls=(-1.5, 0.15), cls_own=1.5, obj=(-3.2, 0.35),
                        obj_lvl=(-4.0, -4.0, 0.0, 0.0), fp16_weights=1.0),
}


def parse_profile(profile):
    """'name' or 'name:key=value,key=value' (experiments) -> dict of calibration targets."""
    name, _, rest = profile.partition(':')
    d = dict(PROFILES[name])
    for kv in rest.split(','):
        if '=' in kv:
            k, v = kv.split('=', 1)
            d[k] = tuple(float(x) for x in v.split('/')) if '/' in v else float(v)
    d['name'] = name
    return d


class _Calibrator:
    def __init__(self, w, no, prof, gen=None):
        self.w, self.no, self.prof, self.gen = w, no, prof, gen
        self.obj_logit = prof['obj']

    def conv(self, x, name, s=1, p=0, act=True):
        w = self.w
        y = F.conv2d(x, w[name + '.weight'], None, stride=s, padding=p)
        if name.startswith('model.33.'):
            no = self.no
            wt = w[name + '.weight'].view(3, no, -1)
            b = w[name + '.bias'].view(3, no)
            yy = y.view(y.shape[0], 3, no, -1)
            for a in range(3):
                for c in range(no):
                    pr = self.prof
                    tgt_std = pr['box_sd'] if c < 4 else (self.obj_logit[1] if c == 4 else pr['cls'][1])
                    tgt_mean = 0.0 if c < 4 else (self.obj_logit[0] if c == 4 else pr['cls'][0])
                    if c == 4 and 'obj_lvl' in pr:
                        # sparse high-confidence candidates: only the coarse levels reach conf >= 0.05, the fine levels
                        # (94 % of the anchors) stay background, as in a real detector's output
                        tgt_mean += pr['obj_lvl'][int(name.rsplit('.', 1)[1])]
                    if c > 4 and 'cls_own' in pr:
                        # anchor slot a votes for class a (mod nc): every class occurs, and no arg-max is a near tie
                        tgt_mean = pr['cls_own'] if (c - 5) == a % (no - 5) else pr['cls'][0]
                    v = yy[:, a, c]
                    sc = tgt_std / float(v.std() + 1e-12)
                    wt[a, c] *= sc
                    b[a, c] = tgt_mean - float(v.mean()) * sc
        else:
            # one scale per layer (not per channel: per-channel statistics of a single calibration image do not
            # transfer to other images and make the network chaotic)
            pr = self.prof
            tgt = pr['res'] if ('.m.' in name and name.endswith('cv2.conv')) else pr['pre']
            w[name + '.weight'] *= tgt / float(y.std() + 1e-12)
            if pr['bias'] is not None:
                # fused-BN biases of a trained detector are of the size of the pre-activations, not 0.1
                b = w[name + '.bias']
                b.copy_(torch.randn(b.shape, generator=self.gen).to(b.device) * (pr['bias'] * tgt))
        y = F.conv2d(x, w[name + '.weight'], w[name + '.bias'], stride=s, padding=p)
        return F.silu(y) if act else y

    def c3(self, x, i, n, shortcut):
        pre = 'model.%d' % i
        a = self.conv(x, pre + '.cv1.conv')
        for j in range(n):
            b = self.conv(self.conv(a, '%s.m.%d.cv1.conv' % (pre, j)), '%s.m.%d.cv2.conv' % (pre, j), p=1)
            a = a + b if shortcut else b
        return self.conv(torch.cat((a, self.conv(x, pre + '.cv2.conv')), 1), pre + '.cv3.conv')

    def run(self, x, n_backbone, n_head):
        nb, nh = n_backbone, n_head
        up = lambda t: F.interpolate(t, scale_factor=2, mode='nearest')
        x0 = self.conv(x, 'model.0.conv', s=2, p=2)
        x2 = self.c3(self.conv(x0, 'model.1.conv', s=2, p=1), 2, nb[0], True)
        x4 = self.c3(self.conv(x2, 'model.3.conv', s=2, p=1), 4, nb[1], True)
        x6 = self.c3(self.conv(x4, 'model.5.conv', s=2, p=1), 6, nb[2], True)
        x8 = self.c3(self.conv(x6, 'model.7.conv', s=2, p=1), 8, nb[3], True)
        x10 = self.c3(self.conv(x8, 'model.9.conv', s=2, p=1), 10, nb[4], True)
        s1 = self.conv(x10, 'model.11.cv1.conv')
        p1 = F.max_pool2d(s1, 5, 1, 2)
        p2 = F.max_pool2d(p1, 5, 1, 2)
        p3 = F.max_pool2d(p2, 5, 1, 2)
        x12 = self.conv(self.conv(torch.cat((s1, p1, p2, p3), 1), 'model.11.cv2.conv'), 'model.12.conv')
        x16 = self.conv(self.c3(torch.cat((up(x12), x8), 1), 15, nh, False), 'model.16.conv')
        x20 = self.conv(self.c3(torch.cat((up(x16), x6), 1), 19, nh, False), 'model.20.conv')
        x23 = self.c3(torch.cat((up(x20), x4), 1), 23, nh, False)
        x26 = self.c3(torch.cat((self.conv(x23, 'model.24.conv', s=2, p=1), x20), 1), 26, nh, False)
        x29 = self.c3(torch.cat((self.conv(x26, 'model.27.conv', s=2, p=1), x16), 1), 29, nh, False)
        x32 = self.c3(torch.cat((self.conv(x29, 'model.30.conv', s=2, p=1), x12), 1), 32, nh, False)
        for l, f in enumerate([x23, x26, x29, x32]):
            self.conv(f, 'model.33.m.%d' % l, act=False)


def _calibration_input(h, w):
    """A synthetic benchmark-style image (index outside every test pool), plainly resized to a stride-64 shape."""
    import cv2
    from .synth import synth_image
    img = synth_image(100000, 1080, 1920)
    img = cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)
    x = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).float() / 255.0
    return x[None]


@torch.no_grad()
def synth_yolov5_weights(ch=YOLOV5X6['ch'], n_backbone=YOLOV5X6['n_backbone'], n_head=4, nc=3, seed=0,
                         calib_hw=(704, 1280), profile='chaotic', device=None):
    """Returns dict name -> float32 CPU tensor.  Deterministic per (arch, seed, calib_hw, device type)."""
    g = torch.Generator().manual_seed(seed)
    w = {}
    prof = parse_profile(profile)
    smooth = float(prof.get('smooth', 0.0))
    coh = float(prof.get('coh', 0.0))
    for name, ci, co, k in _specs(ch, n_backbone, n_head, nc):
        w[name + '.weight'] = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
        if smooth > 0.0 and k > 1:
            # spatially smooth (low-pass) kernels, as trained filters mostly are: a per-(c_out, c_in) gain times a box
            # filter, plus a white remainder.  White rounding noise is then damped relative to the (smooth) signal.
            box = torch.randn(co, ci, 1, 1, generator=g).expand(co, ci, k, k) / (ci * k * k) ** 0.5
            w[name + '.weight'] = (1.0 - smooth ** 2) ** 0.5 * w[name + '.weight'] + smooth * box
        if coh > 0.0 and ci > 3 and not name.startswith('model.33.'):
            w[name + '.weight'] += coh / (ci * k * k) ** 0.5
        if name.startswith('model.33.') and float(prof.get('det_coh', 0.0)) > 0.0:
            # coherent read-out: objectness / class logits follow the local feature energy (all-positive weights plus a
            # white remainder), the way a trained head sums evidence; incoherent rounding noise averages down
            w[name + '.weight'].view(3, nc + 5, -1)[:, 4:, :] += float(prof['det_coh']) / ci ** 0.5
        w[name + '.bias'] = torch.randn(co, generator=g) * 0.1
    # The calibration pass runs on the CPU on purpose: this module is a stand-in for a checkpoint file, and no
    # library kernel (cuDNN) should ever appear on the GPU in a process that is measuring the engine.
    if device is None:
        device = 'cpu'
    if device != 'cpu':
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    threads = torch.get_num_threads()
    try:
        # oneDNN convolutions collapse when oversubscribed (128 threads: 100 GFLOP/s, 32 threads: 1.4 TFLOP/s, measured)
        torch.set_num_threads(max(1, min(threads, 32)))
        wd = {k: v.to(device) for k, v in w.items()}
        _Calibrator(wd, nc + 5, prof, g).run(_calibration_input(*calib_hw).to(device), n_backbone, n_head)
    finally:
        torch.set_num_threads(threads)
    out = {k: v.float().cpu().contiguous() for k, v in wd.items()}
    if prof.get('fp16_weights'):
        # fused weights exactly representable in half precision (yolov5 checkpoints store half-precision weights), so that
        # the engine's fp32 -> fp16 weight conversion is lossless; biases stay fp32
        out = {k: (v.half().float() if k.endswith('.weight') else v) for k, v in out.items()}
    return out
