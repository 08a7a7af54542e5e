"""
ORACLE (test infrastructure only -- never imported by the product path).

Plain-PyTorch fp32 restatement of the YOLOv5-P6 `DetectionModel` forward that the reference executes at
/root/reference/megadetector/detection/pytorch_detector.py:1313 (`self.model(batch_tensor)[0]`) after
`.float().fuse().eval()` (:957).

The module graph lives in the un-vendored dependency `ultralytics-yolov5 == 0.1.1`
(/root/reference/pyproject.toml:70; imports at pytorch_detector.py:333-338); neither it nor MDv5 weights
exist offline.  PARITY UNPINNED for this file: the topology follows the public `yolov5x6.yaml`
 and is checked only through invariants (139.99 M fused parameters at nc=3,
140.73 M at nc=80, 249.49 GMAC at 768x1280).  Everything the GPU engine is compared with is THIS
restatement, run in fp32.

Weight naming follows the fused yolov5 state dict so a real checkpoint maps 1:1:
    model.{i}.conv.{weight,bias}                     plain Conv
    model.{i}.cv{1,2,3}.conv.*, model.{i}.m.{j}.cv{1,2}.conv.*      C3
    model.11.cv{1,2}.conv.*                           SPPF
    model.33.m.{l}.{weight,bias}                      Detect 1x1 heads
"""

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_ANCHORS_PX = [
    [19, 27, 44, 40, 38, 94],
    [96, 68, 86, 152, 180, 137],
    [140, 301, 303, 264, 238, 542],
    [436, 615, 739, 380, 925, 792],
]
STRIDES = [8, 16, 32, 64]


@dataclass
class Arch:
    """YOLOv5-P6 family member.  Defaults = yolov5x6 (depth 1.33, width 1.25) with MegaDetector's nc=3."""
    ch: list = field(default_factory=lambda: [80, 160, 320, 640, 960, 1280])
    n_backbone: list = field(default_factory=lambda: [4, 8, 12, 4, 4])  # C3 repeats at idx 2,4,6,8,10
    n_head: int = 4
    nc: int = 3
    anchors_px: list = field(default_factory=lambda: [list(a) for a in DEFAULT_ANCHORS_PX])

    @property
    def no(self):
        return self.nc + 5


def tiny_arch():
    """Same topology, quarter the width and shallow: used by fast tests."""
    return Arch(ch=[32, 64, 128, 256, 384, 512], n_backbone=[1, 2, 2, 1, 1], n_head=1, nc=3)


def conv_specs(arch):
    """
    Ordered list of every fused conv: (name, c_in, c_out, k, s, p, act).  This order is also the
    canonical execution order used by the engine's plan.
    """
    c = arch.ch
    specs = []

    def conv(name, ci, co, k, s, p=None, act=True):
        specs.append((name, ci, co, k, s, k // 2 if p is None else p, act))

    def c3(i, ci, co, n):
        h = co // 2
        conv('model.%d.cv1.conv' % i, ci, h, 1, 1)
        conv('model.%d.cv2.conv' % i, ci, h, 1, 1)
        for j in range(n):
            conv('model.%d.m.%d.cv1.conv' % (i, j), h, h, 1, 1)
            conv('model.%d.m.%d.cv2.conv' % (i, j), h, h, 3, 1)
        conv('model.%d.cv3.conv' % i, 2 * h, co, 1, 1)

    nb = arch.n_backbone
    conv('model.0.conv', 3, c[0], 6, 2, 2)
    conv('model.1.conv', c[0], c[1], 3, 2)
    c3(2, c[1], c[1], nb[0])
    conv('model.3.conv', c[1], c[2], 3, 2)
    c3(4, c[2], c[2], nb[1])
    conv('model.5.conv', c[2], c[3], 3, 2)
    c3(6, c[3], c[3], nb[2])
    conv('model.7.conv', c[3], c[4], 3, 2)
    c3(8, c[4], c[4], nb[3])
    conv('model.9.conv', c[4], c[5], 3, 2)
    c3(10, c[5], c[5], nb[4])
    conv('model.11.cv1.conv', c[5], c[5] // 2, 1, 1)
    conv('model.11.cv2.conv', c[5] * 2, c[5], 1, 1)
    nh = arch.n_head
    conv('model.12.conv', c[5], c[4], 1, 1)
    c3(15, 2 * c[4], c[4], nh)
    conv('model.16.conv', c[4], c[3], 1, 1)
    c3(19, 2 * c[3], c[3], nh)
    conv('model.20.conv', c[3], c[2], 1, 1)
    c3(23, 2 * c[2], c[2], nh)
    conv('model.24.conv', c[2], c[2], 3, 2)
    c3(26, 2 * c[2], c[3], nh)
    conv('model.27.conv', c[3], c[3], 3, 2)
    c3(29, 2 * c[3], c[4], nh)
    conv('model.30.conv', c[4], c[4], 3, 2)
    c3(32, 2 * c[4], c[5], nh)
    for l, ci in enumerate([c[2], c[3], c[4], c[5]]):
        specs.append(('model.33.m.%d' % l, ci, 3 * arch.no, 1, 1, 0, False))
    return specs


def count_params_and_macs(arch, h, w):
    """Fused parameter count and conv multiply-accumulates for one h x w input."""
    params, macs = 0, 0
    # spatial size per conv: follow the graph
    for name, ci, co, k, s, p, act in conv_specs(arch):
        params += co * ci * k * k + co
    # MACs: walk by stage resolution
    lvl = {0: 2, 1: 4, 2: 4, 3: 8, 4: 8, 5: 16, 6: 16, 7: 32, 8: 32, 9: 64, 10: 64, 11: 64, 12: 64, 15: 32, 16: 32,
           19: 16, 20: 16, 23: 8, 24: 16, 26: 16, 27: 32, 29: 32, 30: 64, 32: 64}
    det_lvl = [8, 16, 32, 64]
    for name, ci, co, k, s, p, act in conv_specs(arch):
        idx = int(name.split('.')[1])
        if idx == 33:
            st = det_lvl[int(name.split('.')[3])]
        else:
            st = lvl[idx]
        macs += (h // st) * (w // st) * co * ci * k * k
    return params, macs


def _conv(x, w, name, s=1, p=0, act=True, trace=None):
    y = F.conv2d(x, w[name + '.weight'], w[name + '.bias'], stride=s, padding=p)
    if act:
        y = F.silu(y)
    if trace is not None:
        trace[name] = y
    return y


def _c3(x, w, i, n, shortcut, trace):
    pre = 'model.%d' % i
    a = _conv(x, w, pre + '.cv1.conv', trace=trace)
    for j in range(n):
        b = _conv(a, w, '%s.m.%d.cv1.conv' % (pre, j), trace=trace)
        b = _conv(b, w, '%s.m.%d.cv2.conv' % (pre, j), p=1, trace=trace)
        a = a + b if shortcut else b
        if trace is not None:
            trace['%s.m.%d' % (pre, j)] = a
    c = _conv(x, w, pre + '.cv2.conv', trace=trace)
    return _conv(torch.cat((a, c), 1), w, pre + '.cv3.conv', trace=trace)


def decode_level(raw, level, arch):
    """
    yolov5 (v7.0 era, = ultralytics-yolov5 0.1.1) `Detect.forward` inference branch for one level.
    raw: [B, na*no, ny, nx] conv output -> [B, na*ny*nx, no] rows ordered (anchor, y, x).
    """
    B, _, ny, nx = raw.shape
    na, no = 3, arch.no
    x = raw.view(B, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
    y = x.sigmoid()
    dev = raw.device
    yv, xv = torch.meshgrid(torch.arange(ny, dtype=torch.float32, device=dev),
                            torch.arange(nx, dtype=torch.float32, device=dev), indexing='ij')
    grid = torch.stack((xv, yv), 2).expand(1, na, ny, nx, 2) - 0.5
    anchor_grid = torch.tensor(arch.anchors_px[level], dtype=torch.float32, device=dev).view(1, na, 1, 1, 2)
    xy = (y[..., 0:2] * 2 + grid) * float(STRIDES[level])
    wh = (y[..., 2:4] * 2) ** 2 * anchor_grid
    out = torch.cat((xy, wh, y[..., 4:]), 4)
    return out.view(B, na * ny * nx, no)


@torch.no_grad()
def forward(x, w, arch=None, trace=None):
    """x: float32 [B,3,H,W] in [0,1].  Returns pred [B, n_anchors, 5+nc] (and fills `trace` if given)."""
    arch = arch or Arch()
    nb, nh = arch.n_backbone, arch.n_head
    x0 = _conv(x, w, 'model.0.conv', s=2, p=2, trace=trace)
    x1 = _conv(x0, w, 'model.1.conv', s=2, p=1, trace=trace)
    x2 = _c3(x1, w, 2, nb[0], True, trace)
    x3 = _conv(x2, w, 'model.3.conv', s=2, p=1, trace=trace)
    x4 = _c3(x3, w, 4, nb[1], True, trace)
    x5 = _conv(x4, w, 'model.5.conv', s=2, p=1, trace=trace)
    x6 = _c3(x5, w, 6, nb[2], True, trace)
    x7 = _conv(x6, w, 'model.7.conv', s=2, p=1, trace=trace)
    x8 = _c3(x7, w, 8, nb[3], True, trace)
    x9 = _conv(x8, w, 'model.9.conv', s=2, p=1, trace=trace)
    x10 = _c3(x9, w, 10, nb[4], True, trace)
    # SPPF
    s1 = _conv(x10, w, 'model.11.cv1.conv', trace=trace)
    p1 = F.max_pool2d(s1, 5, 1, 2)
    p2 = F.max_pool2d(p1, 5, 1, 2)
    p3 = F.max_pool2d(p2, 5, 1, 2)
    x11 = _conv(torch.cat((s1, p1, p2, p3), 1), w, 'model.11.cv2.conv', trace=trace)
    # head
    x12 = _conv(x11, w, 'model.12.conv', trace=trace)
    x15 = _c3(torch.cat((F.interpolate(x12, scale_factor=2, mode='nearest'), x8), 1), w, 15, nh, False, trace)
    x16 = _conv(x15, w, 'model.16.conv', trace=trace)
    x19 = _c3(torch.cat((F.interpolate(x16, scale_factor=2, mode='nearest'), x6), 1), w, 19, nh, False, trace)
    x20 = _conv(x19, w, 'model.20.conv', trace=trace)
    x23 = _c3(torch.cat((F.interpolate(x20, scale_factor=2, mode='nearest'), x4), 1), w, 23, nh, False, trace)
    x24 = _conv(x23, w, 'model.24.conv', s=2, p=1, trace=trace)
    x26 = _c3(torch.cat((x24, x20), 1), w, 26, nh, False, trace)
    x27 = _conv(x26, w, 'model.27.conv', s=2, p=1, trace=trace)
    x29 = _c3(torch.cat((x27, x16), 1), w, 29, nh, False, trace)
    x30 = _conv(x29, w, 'model.30.conv', s=2, p=1, trace=trace)
    x32 = _c3(torch.cat((x30, x12), 1), w, 32, nh, False, trace)
    z = []
    for l, f in enumerate([x23, x26, x29, x32]):
        raw = _conv(f, w, 'model.33.m.%d' % l, act=False, trace=trace)
        z.append(decode_level(raw, l, arch))
    return torch.cat(z, 1)


# ---------------------------------------------------------------------------------------------------------------
# Test-time augmentation: `model(x, augment=True)` (pytorch_detector.py:1313 with augment) -> yolov5 models/yolo.py
# `DetectionModel._forward_augment` / `_descale_pred` / `_clip_augmented` and utils/torch_utils.py `scale_img`
# (ultralytics-yolov5 0.1.1, absent offline: restated from the public source, PARITY UNPINNED like the rest of this file)
# ---------------------------------------------------------------------------------------------------------------
TTA_SCALES = (1, 0.83, 0.67)
TTA_FLIPS = (None, 3, None)  # 3 = left-right


def scale_img(img, ratio=1.0, gs=64):
    """yolov5 scale_img(img, ratio, same_shape=False, gs): bilinear resize, then pad right/bottom to a gs multiple with 0.447."""
    import math
    if ratio == 1.0:
        return img
    h, w = img.shape[2:]
    s = (int(h * ratio), int(w * ratio))
    img = F.interpolate(img, size=s, mode='bilinear', align_corners=False)
    h, w = (math.ceil(x * ratio / gs) * gs for x in (h, w))
    return F.pad(img, [0, w - s[1], 0, h - s[0]], value=0.447)


def _descale_pred(p, flips, scale, img_size):
    p = p.clone()
    p[..., :4] /= scale
    if flips == 2:
        p[..., 1] = img_size[0] - p[..., 1]
    elif flips == 3:
        p[..., 0] = img_size[1] - p[..., 0]
    return p


def _clip_augmented(y, nl=4):
    g = sum(4 ** x for x in range(nl))
    e = 1
    i = (y[0].shape[1] // g) * sum(4 ** x for x in range(e))
    y[0] = y[0][:, :-i]
    i = (y[-1].shape[1] // g) * sum(4 ** (nl - 1 - x) for x in range(e))
    y[-1] = y[-1][:, i:]
    return y


@torch.no_grad()
def forward_augment(x, w, arch=None):
    """pred of `model(x, augment=True)[0]`: the three scaled / flipped passes, de-scaled, clipped and concatenated."""
    arch = arch or Arch()
    img_size = x.shape[-2:]
    y = []
    for si, fi in zip(TTA_SCALES, TTA_FLIPS):
        xi = scale_img(x.flip(fi) if fi else x, si, gs=STRIDES[-1])
        yi = forward(xi, w, arch)
        y.append(_descale_pred(yi, fi, si, img_size))
    y = _clip_augmented(y)
    return torch.cat(y, 1)
