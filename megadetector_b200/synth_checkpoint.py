"""
Writes a synthetic weight set (megadetector_b200/synth_weights.py) as a yolov5-format `.pt` FILE, so that the reference's own
drivers (`run_detector.load_detector` picks the detector class by the `.pt` extension, run_detector.py:643; MDv5a itself is
`torch.save({'model': DetectionModel, ...})`, loaded at pytorch_detector.py:929) can be pointed at it.  No MDv5 checkpoint
exists offline; this is the stand-in "model file" of BASELINE.json's config 1.

The pickle refers to classes `models.yolo.DetectionModel`, `models.yolo.Detect`, `models.common.Conv/C3/Bottleneck/SPPF/
Concat` like a real yolov5 checkpoint; `checkpoint.load_yolov5_checkpoint` reads it back through placeholder classes.
Convolutions are stored already fused (Conv2d with bias, no BatchNorm), in fp32.
"""
import sys
import types

import torch
import torch.nn as nn

from . import checkpoint


def _modules():
    common = types.ModuleType('models.common')
    yolo = types.ModuleType('models.yolo')
    models = types.ModuleType('models')

    class Conv(nn.Module):
        def __init__(self, w, b, s=1):
            super().__init__()
            co, ci, k, _ = w.shape
            self.conv = nn.Conv2d(ci, co, k, s, k // 2, bias=True)
            self.conv.weight.data.copy_(w)
            self.conv.bias.data.copy_(b)
            self.act = nn.SiLU()

    class Bottleneck(nn.Module):
        def __init__(self, cv1, cv2, add):
            super().__init__()
            self.cv1, self.cv2, self.add = cv1, cv2, add

    class C3(nn.Module):
        def __init__(self, cv1, cv2, cv3, m):
            super().__init__()
            self.cv1, self.cv2, self.cv3 = cv1, cv2, cv3
            self.m = nn.Sequential(*m)

    class SPPF(nn.Module):
        def __init__(self, cv1, cv2):
            super().__init__()
            self.cv1, self.cv2 = cv1, cv2
            self.m = nn.MaxPool2d(5, 1, 2)

    class Concat(nn.Module):
        pass

    class Detect(nn.Module):
        def __init__(self, nc, anchors_px, heads):
            super().__init__()
            self.nc, self.no, self.nl, self.na = nc, nc + 5, 4, 3
            stride = torch.tensor(checkpoint.STRIDES)
            self.register_buffer('anchors', torch.tensor(anchors_px).float().view(4, 3, 2) / stride.view(4, 1, 1))
            self.stride = stride
            self.m = nn.ModuleList()
            for w, b in heads:
                c = nn.Conv2d(w.shape[1], w.shape[0], 1)
                c.weight.data.copy_(w)
                c.bias.data.copy_(b)
                self.m.append(c)

    class DetectionModel(nn.Module):
        def __init__(self, layers):
            super().__init__()
            self.model = nn.Sequential(*layers)

    for cls in (Conv, Bottleneck, C3, SPPF, Concat):
        cls.__module__ = 'models.common'
        cls.__qualname__ = cls.__name__
        setattr(common, cls.__name__, cls)
    for cls in (Detect, DetectionModel):
        cls.__module__ = 'models.yolo'
        cls.__qualname__ = cls.__name__
        setattr(yolo, cls.__name__, cls)
    return {'models': models, 'models.common': common, 'models.yolo': yolo}


def write_synthetic_checkpoint(spec, path):
    """spec: a `synthetic://...` model string; path: the `.pt` file to write.  Returns the arch dict."""
    from . import _capi
    w, arch = checkpoint.load_model(spec)
    anchors = arch.get('anchors_px') or _capi.DEFAULT_ANCHORS_PX
    mods = _modules()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        c, y = mods['models.common'], mods['models.yolo']

        def conv(name, s=1):
            return c.Conv(w[name + '.conv.weight'], w[name + '.conv.bias'], s)

        def c3(i, n, add):
            pre = 'model.%d' % i
            m = [c.Bottleneck(conv('%s.m.%d.cv1' % (pre, j)), conv('%s.m.%d.cv2' % (pre, j)), add) for j in range(n)]
            return c.C3(conv(pre + '.cv1'), conv(pre + '.cv2'), conv(pre + '.cv3'), m)

        nb, nh = arch['n_backbone'], arch['n_head']
        up = lambda: nn.Upsample(scale_factor=2, mode='nearest')
        L = [conv('model.0', 2), conv('model.1', 2), c3(2, nb[0], True), conv('model.3', 2), c3(4, nb[1], True),
             conv('model.5', 2), c3(6, nb[2], True), conv('model.7', 2), c3(8, nb[3], True), conv('model.9', 2),
             c3(10, nb[4], True), c.SPPF(conv('model.11.cv1'), conv('model.11.cv2')), conv('model.12'), up(), c.Concat(),
             c3(15, nh, False), conv('model.16'), up(), c.Concat(), c3(19, nh, False), conv('model.20'), up(), c.Concat(),
             c3(23, nh, False), conv('model.24', 2), c.Concat(), c3(26, nh, False), conv('model.27', 2), c.Concat(),
             c3(29, nh, False), conv('model.30', 2), c.Concat(), c3(32, nh, False),
             y.Detect(arch['nc'], anchors, [(w['model.33.m.%d.weight' % l], w['model.33.m.%d.bias' % l]) for l in range(4)])]
        model = y.DetectionModel(L)
        torch.save({'model': model, 'epoch': -1, 'synthetic_spec': spec}, path)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return arch
