"""CPU: loading a yolov5-style `.pt` without the yolov5 package (placeholder unpickling + Conv/BN folding)."""
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from megadetector_b200 import checkpoint, synth_weights


def _fake_yolov5_modules():
    common = types.ModuleType('models.common')
    yolo = types.ModuleType('models.yolo')
    models = types.ModuleType('models')

    class Conv(nn.Module):
        def __init__(self, c1, c2, k=1, s=1, p=None):
            super().__init__()
            self.conv = nn.Conv2d(c1, c2, k, s, k // 2 if p is None else p, bias=False)
            self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
            self.act = nn.SiLU()

    class Bottleneck(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.cv1 = Conv(c, c, 1)
            self.cv2 = Conv(c, c, 3)

    class C3(nn.Module):
        def __init__(self, c1, c2, n):
            super().__init__()
            h = c2 // 2
            self.cv1 = Conv(c1, h)
            self.cv2 = Conv(c1, h)
            self.cv3 = Conv(2 * h, c2)
            self.m = nn.Sequential(*(Bottleneck(h) for _ in range(n)))

    class SPPF(nn.Module):
        def __init__(self, c1, c2):
            super().__init__()
            self.cv1 = Conv(c1, c1 // 2)
            self.cv2 = Conv(c1 * 2, c2)
            self.m = nn.MaxPool2d(5, 1, 2)

    class Concat(nn.Module):
        pass

    class Detect(nn.Module):
        def __init__(self, nc, anchors, ch):
            super().__init__()
            self.nc, self.no, self.nl, self.na = nc, nc + 5, 4, 3
            self.register_buffer('anchors', torch.tensor(anchors).float().view(4, 3, 2))
            self.m = nn.ModuleList(nn.Conv2d(c, 3 * (nc + 5), 1) for c in ch)
            self.stride = torch.tensor([8., 16., 32., 64.])

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


def test_load_checkpoint_without_yolov5_package(tmp_path):
    mods = _fake_yolov5_modules()
    sys.modules.update(mods)
    try:
        c, y = mods['models.common'], mods['models.yolo']
        ch = (32, 64, 128, 256, 384, 512)
        nb, nh, nc = (1, 2, 2, 1, 1), 1, 3
        anchors_units = (np.array(synth_weights_default_anchors()) / np.repeat([[8.], [16.], [32.], [64.]], 6, 1)).tolist()
        L = [c.Conv(3, ch[0], 6, 2, 2), c.Conv(ch[0], ch[1], 3, 2), c.C3(ch[1], ch[1], nb[0]), c.Conv(ch[1], ch[2], 3, 2),
             c.C3(ch[2], ch[2], nb[1]), c.Conv(ch[2], ch[3], 3, 2), c.C3(ch[3], ch[3], nb[2]), c.Conv(ch[3], ch[4], 3, 2),
             c.C3(ch[4], ch[4], nb[3]), c.Conv(ch[4], ch[5], 3, 2), c.C3(ch[5], ch[5], nb[4]), c.SPPF(ch[5], ch[5]),
             c.Conv(ch[5], ch[4], 1), nn.Upsample(scale_factor=2), c.Concat(), c.C3(2 * ch[4], ch[4], nh),
             c.Conv(ch[4], ch[3], 1), nn.Upsample(scale_factor=2), c.Concat(), c.C3(2 * ch[3], ch[3], nh),
             c.Conv(ch[3], ch[2], 1), nn.Upsample(scale_factor=2), c.Concat(), c.C3(2 * ch[2], ch[2], nh),
             c.Conv(ch[2], ch[2], 3, 2), c.Concat(), c.C3(2 * ch[2], ch[3], nh), c.Conv(ch[3], ch[3], 3, 2), c.Concat(),
             c.C3(2 * ch[3], ch[4], nh), c.Conv(ch[4], ch[4], 3, 2), c.Concat(), c.C3(2 * ch[4], ch[5], nh),
             y.Detect(nc, anchors_units, (ch[2], ch[3], ch[4], ch[5]))]
        torch.manual_seed(0)
        model = y.DetectionModel(L)
        for m in model.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.5)
                m.running_var.uniform_(0.5, 2.0)
                m.weight.data.uniform_(0.5, 1.5)
                m.bias.data.normal_(0, 0.2)
        # expected fused tensors, the way yolov5 fuse_conv_and_bn computes them
        expect = {}
        for name, m in model.named_modules():
            if type(m).__name__ == 'Conv':
                w_bn = torch.diag(m.bn.weight.div(torch.sqrt(m.bn.eps + m.bn.running_var)))
                fw = torch.mm(w_bn, m.conv.weight.detach().view(m.conv.out_channels, -1)).view(m.conv.weight.shape)
                fb = torch.mm(w_bn, torch.zeros(m.conv.out_channels, 1)).reshape(-1) + \
                    (m.bn.bias - m.bn.weight.mul(m.bn.running_mean).div(torch.sqrt(m.bn.running_var + m.bn.eps)))
                expect[name + '.conv.weight'] = fw.detach()
                expect[name + '.conv.bias'] = fb.detach()
        path = str(tmp_path / 'fake_md.pt')
        torch.save({'model': model.half(), 'epoch': -1}, path)
    finally:
        for k in mods:
            sys.modules.pop(k, None)
    assert 'models' not in sys.modules
    w, arch = checkpoint.load_yolov5_checkpoint(path)
    assert arch['ch'] == ch and arch['n_backbone'] == nb and arch['n_head'] == nh and arch['nc'] == nc
    assert np.allclose(np.array(arch['anchors_px']), np.array(synth_weights_default_anchors()))
    names = [s[0] for s in synth_weights._specs(ch, nb, nh, nc)]
    for n in names:
        assert n + '.weight' in w and n + '.bias' in w, n
    # the checkpoint is stored in fp16 (like MDv5); folding happens after .float(), like the reference (:957)
    for k, v in expect.items():
        assert w[k].dtype == torch.float32 and w[k].shape == v.shape
        assert torch.allclose(w[k], v, rtol=2e-3, atol=2e-3), k
    assert checkpoint.read_metadata_from_megadetector_model_file(path) is None


def synth_weights_default_anchors():
    from megadetector_b200._capi import DEFAULT_ANCHORS_PX
    return [list(map(float, a)) for a in DEFAULT_ANCHORS_PX]


def test_synthetic_specs():
    assert checkpoint.is_synthetic('synthetic://yolov5x6?seed=3')
    arch, seed, _ = checkpoint._parse_synthetic('synthetic://yolov5x6?seed=3')
    assert seed == 3 and tuple(arch['ch']) == (80, 160, 320, 640, 960, 1280)
    arch, seed, _ = checkpoint._parse_synthetic('synthetic://mdv5b')
    assert seed == 1
    assert len(synth_weights._specs(**{k: arch[k] for k in ('ch', 'n_backbone', 'n_head', 'nc')})) == 163
