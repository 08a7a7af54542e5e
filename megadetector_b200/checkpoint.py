"""
Model-file handling: replaces `PTDetector._load_model` (megadetector/detection/pytorch_detector.py:914-959), which
does `torch.load(model.pt)['model'].float().fuse().eval()` with the yolov5 package importable.

We do not need (or have) the yolov5 package: the checkpoint is unpickled with placeholder classes standing in
for `models.yolo.*` / `models.common.*`, the tensors are read from the module tree, Conv+BatchNorm pairs are
folded exactly as yolov5's `fuse_conv_and_bn` does (fp32, eps taken from the pickled BN module), and the
architecture (widths, C3 depths, nc, anchors) is recovered from the tensor shapes.

A second source exists for benchmarks and tests, because no MDv5 checkpoint is available offline:
    "synthetic://yolov5x6?seed=0"      calibrated random-init weights (megadetector_b200/synth_weights.py)
    "synthetic://tiny?seed=0"          same topology, quarter width (fast tests)
"""

import io
import pickle
import zipfile

import torch

from . import synth_weights

P6_DETECT_INDEX = 33
STRIDES = (8.0, 16.0, 32.0, 64.0)


def is_synthetic(model_path):
    return isinstance(model_path, str) and model_path.startswith('synthetic://')


def _parse_synthetic(model_path):
    rest = model_path[len('synthetic://'):]
    name, _, query = rest.partition('?')
    opts = dict(kv.split('=', 1) for kv in query.split('&') if '=' in kv)
    arch = dict(synth_weights.YOLOV5X6 if name in ('yolov5x6', 'mdv5a', 'mdv5b') else synth_weights.TINY_P6)
    seed = int(opts.get('seed', 1 if name == 'mdv5b' else 0))
    return arch, seed, opts.get('profile', 'chaotic')


class _Placeholder(torch.nn.Module):
    """Stands in for any yolov5 class found in the pickle; nn.Module.__setstate__ restores the tensors."""

    def __init__(self, *args, **kwargs):
        super().__init__()


class _ShimUnpickler(pickle.Unpickler):
    _cache = {}

    def find_class(self, module, name):
        root = module.split('.')[0]
        if root in ('models', 'utils', 'yolov5', 'ultralytics'):
            key = module + '.' + name
            if key not in self._cache:
                self._cache[key] = type(name, (_Placeholder,), {'__module__': module})
            return self._cache[key]
        return super().find_class(module, name)


class _ShimPickleModule:
    Unpickler = _ShimUnpickler
    __name__ = 'pickle'

    @staticmethod
    def load(f, **kwargs):
        return _ShimUnpickler(f, **kwargs).load()

    Pickler = pickle.Pickler
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    loads = staticmethod(pickle.loads)


def _fold_bn(conv_w, conv_b, bn_w, bn_b, mean, var, eps):
    """yolov5 utils/torch_utils.py::fuse_conv_and_bn, elementwise form (identical rounding: one fp32 multiply)."""
    scale = bn_w.float().div(torch.sqrt(eps + var.float()))
    w = conv_w.float() * scale.view(-1, 1, 1, 1)
    b0 = torch.zeros_like(scale) if conv_b is None else conv_b.float()
    b = scale * b0 + (bn_b.float() - bn_w.float().mul(mean.float()).div(torch.sqrt(var.float() + eps)))
    return w.contiguous(), b.contiguous()


def _arch_from_weights(w, anchors_px):
    def co(name):
        return int(w[name + '.weight'].shape[0])

    ch = (co('model.0.conv'), co('model.1.conv'), co('model.3.conv'), co('model.5.conv'), co('model.7.conv'),
          co('model.9.conv'))

    def depth(i):
        j = 0
        while 'model.%d.m.%d.cv1.conv.weight' % (i, j) in w:
            j += 1
        return j

    n_backbone = tuple(depth(i) for i in (2, 4, 6, 8, 10))
    n_head = depth(15)
    for i in (19, 23, 26, 29, 32):
        if depth(i) != n_head:
            raise ValueError('unsupported checkpoint: head C3 depths differ')
    no = int(w['model.%d.m.0.weight' % P6_DETECT_INDEX].shape[0]) // 3
    return dict(ch=ch, n_backbone=n_backbone, n_head=n_head, nc=no - 5, anchors_px=anchors_px)


def load_yolov5_checkpoint(path):
    """Returns (weights dict name -> fp32 tensor [fused], arch dict) for a YOLOv5-P6 `.pt` file."""
    ckpt = torch.load(path, map_location='cpu', weights_only=False, pickle_module=_ShimPickleModule)
    model = ckpt['model'] if isinstance(ckpt, dict) and 'model' in ckpt else ckpt
    sd = {k: v for k, v in model.state_dict().items()}
    mods = dict(model.named_modules())
    w = {}
    for name in list(sd.keys()):
        if not name.endswith('.conv.weight'):
            continue
        base = name[:-len('.conv.weight')]  # e.g. model.2.cv1
        bn = base + '.bn'
        conv_b = sd.get(base + '.conv.bias')
        if bn + '.weight' in sd:
            eps = float(getattr(mods.get(bn), 'eps', 1e-3))
            fw, fb = _fold_bn(sd[name], conv_b, sd[bn + '.weight'], sd[bn + '.bias'], sd[bn + '.running_mean'],
                              sd[bn + '.running_var'], eps)
        else:  # already fused
            fw, fb = sd[name].float().contiguous(), (conv_b.float() if conv_b is not None else
                                                     torch.zeros(sd[name].shape[0])).contiguous()
        w[base + '.conv.weight'] = fw
        w[base + '.conv.bias'] = fb
    det = 'model.%d' % P6_DETECT_INDEX
    if det + '.m.0.weight' not in sd:
        raise ValueError('unsupported checkpoint: expected a YOLOv5-P6 model with Detect at index {}'.format(P6_DETECT_INDEX))
    for l in range(4):
        w['%s.m.%d.weight' % (det, l)] = sd['%s.m.%d.weight' % (det, l)].float().contiguous()
        w['%s.m.%d.bias' % (det, l)] = sd['%s.m.%d.bias' % (det, l)].float().contiguous()
    # Detect.anchors is stored in units of the level stride; anchor_grid = anchors * stride
    anchors = sd[det + '.anchors'].float()
    stride = sd.get(det + '.stride')
    if stride is None:
        stride = getattr(mods.get(det), 'stride', None)
    stride = torch.tensor(STRIDES) if stride is None else torch.as_tensor(stride).float()
    anchors_px = (anchors * stride.view(-1, 1, 1)).reshape(4, 6).tolist()
    return w, _arch_from_weights(w, anchors_px)


def load_model(model_path):
    """(weights, arch) for a `.pt` path or a `synthetic://` spec."""
    if is_synthetic(model_path):
        arch, seed, profile = _parse_synthetic(model_path)
        w = synth_weights.synth_yolov5_weights(arch['ch'], arch['n_backbone'], arch['n_head'], arch['nc'], seed=seed,
                                               profile=profile)
        arch['anchors_px'] = None
        return w, arch
    return load_yolov5_checkpoint(model_path)


def read_metadata_from_megadetector_model_file(model_file):
    """megadetector_info.json inside the .pt zip (pytorch_detector.py:674-728); None when absent."""
    import json
    if is_synthetic(model_file):
        return None
    try:
        with zipfile.ZipFile(model_file, 'r') as z:
            names = [n for n in z.namelist() if n.endswith('megadetector_info.json')]
            if not names:
                return None
            with z.open(names[0]) as f:
                return json.load(io.TextIOWrapper(f, encoding='utf-8'))
    except Exception:
        return None
