"""Torch-free, seconds-long GPU check of the two halo main loops (stride 1 single box, stride 2 parity boxes) against a numpy
convolution of the same fp16-rounded operands: python tools/quick_halo_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from megadetector_b200 import _capi


def ref_conv(x16, w, b, s):
    x = np.pad(x16.astype(np.float32), ((0, 0), (1, 1), (1, 1), (0, 0)))
    wt = w.astype(np.float16).astype(np.float32)
    n, hp, wp, _ = x.shape
    ho, wo = (hp - 3) // s + 1, (wp - 3) // s + 1
    y = np.zeros((n, ho, wo, w.shape[0]), np.float64)
    for kh in range(3):
        for kw in range(3):
            patch = x[:, kh:kh + s * (ho - 1) + 1:s, kw:kw + s * (wo - 1) + 1:s, :]
            y += np.einsum('nhwc,oc->nhwo', patch, wt[:, :, kh, kw], optimize=True)
    y += b
    return (y / (1.0 + np.exp(-y))).astype(np.float32)


e = _capi.Engine(_capi.make_config(max_batch=2))
rng = np.random.default_rng(0)
ok = True
for name, h, w_, cin, cout, s in (('s1 single box', 32, 48, 128, 64, 1), ('s2 parity boxes', 64, 96, 128, 64, 2),
                                  ('s1 tail chunk', 32, 48, 80, 80, 1)):
    x = rng.standard_normal((2, h, w_, cin)).astype(np.float16)
    wt = (rng.standard_normal((cout, cin, 3, 3)) / np.sqrt(cin * 9)).astype(np.float32)
    b = (rng.standard_normal(cout) * 0.5).astype(np.float32)
    y, _ = e.op_conv2d(x, wt, b, 3, s, 1, act=True)
    r = ref_conv(x, wt, b, s)
    err = float(np.max(np.abs(y.astype(np.float32) - r) / np.maximum(1.0, np.abs(r))))
    print('%-16s max err / tol = %.3f' % (name, err / 3e-3), flush=True)
    ok = ok and err <= 3e-3
print('HALO CHECK', 'OK' if ok else 'FAILED')
sys.exit(0 if ok else 1)
