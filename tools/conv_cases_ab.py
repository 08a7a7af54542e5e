"""conv_cases.py with the package root as first argument: python tools/conv_cases_ab.py <tree> <cases> <iters> (same-box A/B of two trees)."""
import os
import sys

ROOT = os.path.abspath(sys.argv.pop(1))  # which tree's megadetector_b200 to load
sys.path.insert(0, ROOT)
import numpy as np

from megadetector_b200 import _capi

CASES = {
    'c1x1_160': (16, 96, 160, 160, 160, 1, 1, 0, False),
    'c1x1_320': (16, 48, 80, 320, 320, 1, 1, 0, False),
    'c3x3_80': (16, 192, 320, 80, 80, 3, 1, 1, True),
    'c3x3_160': (16, 96, 160, 160, 160, 3, 1, 1, True),
    'c3x3_320': (16, 48, 80, 320, 320, 3, 1, 1, True),
    'c3x3_480': (16, 24, 40, 480, 480, 3, 1, 1, True),
    'stem': (16, 384, 640, 16, 80, 3, 1, 1, False),
    's2_80_160': (16, 384, 640, 80, 160, 3, 2, 1, False),
    'c1x1_160_noact': (16, 96, 160, 160, 160, 1, 1, 0, False),
    'c1x1_80': (16, 192, 320, 80, 80, 1, 1, 0, False),
    'c1x1_640': (16, 48, 80, 640, 640, 1, 1, 0, False),
    'p2_c64': (16, 192, 320, 64, 80, 3, 1, 1, False),
    'p2_c80': (16, 192, 320, 80, 80, 3, 1, 1, False),
    'p2_c128': (16, 192, 320, 128, 80, 3, 1, 1, False),
    'stem_c64': (16, 384, 640, 64, 80, 3, 1, 1, False),
    'b64_c1x1_80': (64, 192, 320, 80, 80, 1, 1, 0, False),
    'b64_c1x1_160': (64, 96, 160, 160, 160, 1, 1, 0, False),
    'b64_c1x1_160_80': (64, 192, 320, 160, 80, 1, 1, 0, False),
    'b64_c3x3_80': (64, 192, 320, 80, 80, 3, 1, 1, True),
    'b64_c1x1_320': (64, 48, 80, 320, 320, 1, 1, 0, False),
    'b64_c3x3_160': (64, 96, 160, 160, 160, 3, 1, 1, True),
    'b64_c3x3_320': (64, 48, 80, 320, 320, 3, 1, 1, True),
    's2_640_640': (64, 48, 80, 640, 640, 3, 2, 1, False),
    'c1x1_1280_640': (64, 48, 80, 1280, 640, 1, 1, 0, False),
    'b64_s2_80_160': (64, 384, 640, 80, 160, 3, 2, 1, False),
    'b64_s2_160_320': (64, 192, 320, 160, 320, 3, 2, 1, False),
    'b64_s2_320_640': (64, 96, 160, 320, 640, 3, 2, 1, False),
    'b64_s2_320_320': (64, 96, 160, 320, 320, 3, 2, 1, False),
}
names = sys.argv[1].split(',') if len(sys.argv) > 1 else list(CASES)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
e = _capi.Engine(_capi.make_config(max_batch=64))
rng = np.random.default_rng(0)
for nm in names:
    n, h, w, ci, co, k, s, p, res = CASES[nm]
    x = rng.standard_normal((n, h, w, ci)).astype(np.float16)
    wt = (rng.standard_normal((co, ci, k, k)) / np.sqrt(ci * k * k)).astype(np.float32)
    b = np.zeros(co, np.float32)
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    r = rng.standard_normal((n, ho, wo, co)).astype(np.float16) if res else None
    _, ms = e.op_conv2d(x, wt, b, k, s, p, act=not nm.endswith('_noact'), residual=r, iters=iters)
    fl = 2.0 * n * ho * wo * co * ci * k * k
    byt = 2.0 * n * (h * w * ci + ho * wo * co * (2 if res else 1))
    print('%-10s %8.4f ms  %7.1f TFLOP/s  %6.0f GB/s (in+out%s)' % (nm, ms, fl / ms / 1e9, byt / ms / 1e6, '+res' if res else ''), flush=True)
