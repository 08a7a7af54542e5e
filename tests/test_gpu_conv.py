"""
GPU parity of K2 (tcgen05 implicit-GEMM convolution) through the C ABI (`mdb_op_conv2d`) against a plain
PyTorch fp32 convolution of the same fp16-rounded operands.  Tolerance: the kernel accumulates in fp32 and
rounds the result to fp16 once, so |err| <= 2^-10 * |y| + accumulation-order noise; we allow
3e-3 * max(1, |y|) absolute.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # name,             n,  h,  w,  cin, cout, k, s, p, act, res,  fp32out
    ('1x1_c64_single', 1, 8, 16, 64, 64, 1, 1, 0, 1, False, False),
    ('1x1_c128_2chunk', 1, 8, 16, 128, 64, 1, 1, 0, 1, False, False),
    ('1x1_c32_sw64', 1, 8, 16, 32, 32, 1, 1, 0, 0, False, False),
    ('1x1_c16_sw32', 1, 8, 16, 16, 16, 1, 1, 0, 0, False, False),
    ('1x1_c80', 2, 24, 40, 80, 80, 1, 1, 0, 1, False, False),
    ('1x1_c160_res', 2, 24, 40, 160, 160, 1, 1, 0, 1, True, False),
    ('1x1_ragged_m', 1, 13, 7, 64, 96, 1, 1, 0, 1, False, False),
    ('3x3_c64', 1, 16, 16, 64, 64, 3, 1, 1, 1, False, False),
    ('3x3_c64_noact', 1, 16, 16, 64, 64, 3, 1, 1, 0, False, False),
    ('3x3_p6_shape', 2, 12, 20, 64, 160, 3, 1, 1, 1, True, False),
    ('3x3_c80_sw32', 1, 24, 40, 80, 80, 3, 1, 1, 1, True, False),
    ('3x3_c480_sw64', 1, 24, 40, 480, 480, 3, 1, 1, 1, True, False),
    ('3x3_s2_c64', 1, 32, 32, 64, 128, 3, 2, 1, 1, False, False),
    ('3x3_s2_c160', 2, 48, 80, 160, 320, 3, 2, 1, 1, False, False),
    ('stem_like_c16', 1, 96, 160, 16, 80, 3, 1, 1, 1, False, False),
    ('detect_c320_fp32', 2, 24, 40, 320, 24, 1, 1, 0, 0, False, True),
    ('3x3_c320_real', 4, 48, 80, 320, 320, 3, 1, 1, 1, True, False),
    ('1x1_c1280_bn256', 2, 12, 20, 1280, 1280, 1, 1, 0, 1, False, False),
    # several tiles per CTA pair (persistent loop, accumulator ring phases, rotating epilogue chunk assignment)
    ('1x1_c80_many_tiles', 4, 96, 160, 80, 80, 1, 1, 0, 1, False, False),
    ('1x1_c160_many_tiles', 4, 96, 160, 160, 160, 1, 1, 0, 1, True, False),
    ('1x1_c320_two_ntiles', 4, 48, 80, 320, 320, 1, 1, 0, 1, False, False),
    # row-halo main loop from 80 channels (two chunks per tap, the second mostly zero-filled), filter-row weight stages
    ('3x3_c80_halo_many', 4, 96, 160, 80, 80, 3, 1, 1, 1, True, False),
    ('3x3_c160_halo_many', 4, 48, 80, 160, 160, 3, 1, 1, 1, True, False),
    # stride-2 halo main loop: four parity-view boxes per chunk (8|9 px x 16|17 rows), taps addressed by row / pixel offsets
    ('3x3_s2_c80_halo', 2, 64, 96, 80, 160, 3, 2, 1, 1, False, False),
    ('3x3_s2_c160_halo', 2, 96, 160, 160, 320, 3, 2, 1, 1, False, False),
    ('3x3_s2_c320_halo_many', 4, 96, 160, 320, 640, 3, 2, 1, 1, False, False),
    ('3x3_s2_halo_ragged', 1, 60, 80, 128, 64, 3, 2, 1, 1, False, False),
]


@pytest.fixture(scope='module')
def engine():
    from megadetector_b200 import _capi
    e = _capi.Engine(_capi.make_config(max_batch=4))
    yield e
    e.close()


def reference(x16, w, b, k, s, p, act, res16, fp32out):
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda')
    x = torch.from_numpy(x16.astype(np.float32)).to(dev).permute(0, 3, 1, 2).contiguous()
    wt = torch.from_numpy(w.astype(np.float16).astype(np.float32)).to(dev)
    y = F.conv2d(x.double(), wt.double(), torch.from_numpy(b).to(dev).double(), stride=s, padding=p)
    if act:
        y = F.silu(y)
    if res16 is not None:
        y = y + torch.from_numpy(res16.astype(np.float32)).to(dev).permute(0, 3, 1, 2).double()
    return y.permute(0, 2, 3, 1).float().cpu().numpy()


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_conv_matches_fp32_reference(engine, case):
    name, n, h, w_, cin, cout, k, s, p, act, use_res, fp32out = case
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    x = (rng.standard_normal((n, h, w_, cin)) * 1.0).astype(np.float16)
    wt = (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
    b = (rng.standard_normal(cout) * 0.5).astype(np.float32)
    ho, wo = (h + 2 * p - k) // s + 1, (w_ + 2 * p - k) // s + 1
    res = (rng.standard_normal((n, ho, wo, cout))).astype(np.float16) if use_res else None
    y, _ = engine.op_conv2d(x, wt, b, k, s, p, act=bool(act), residual=res, out_fp32=fp32out)
    ref = reference(x, wt, b, k, s, p, act, res, fp32out)
    y32 = y.astype(np.float32)
    assert np.isfinite(y32).all(), '{}: non-finite output ({} bad)'.format(name, int((~np.isfinite(y32)).sum()))
    err = np.abs(y32 - ref)
    tol = (3e-3 if not fp32out else 2e-4) * np.maximum(1.0, np.abs(ref))
    bad = err > tol
    if bad.any():
        idx = np.argwhere(bad)
        first = idx[0]
        # where are the bad entries? (by pixel row / channel) -- this is the remote-debug signal
        per_c = bad.reshape(-1, cout).sum(0)
        per_pix = bad.reshape(n, ho * wo, cout).sum(2)
        msg = ('{}: {} / {} mismatches, max err {:.4g}; first at {} got {:.5g} want {:.5g}; '
               'bad channels {}..{} ({} of {}); bad pixels per image {}; corr {:.4f}; sample got {} want {}').format(
            name, int(bad.sum()), bad.size, float(err.max()), first.tolist(), float(y32[tuple(first)]),
            float(ref[tuple(first)]), int(np.nonzero(per_c)[0].min()), int(np.nonzero(per_c)[0].max()),
            int((per_c > 0).sum()), cout, (per_pix > 0).sum(1).tolist(),
            float(np.corrcoef(y32.ravel(), ref.ravel())[0, 1]), y32.ravel()[:6].tolist(), ref.ravel()[:6].tolist())
        raise AssertionError(msg)


def test_conv_throughput_is_reported(engine):
    # not a parity test: prints the achieved tensor rate of one real MDv5 layer shape (3x3 320->320 @ 48x80, bs 16)
    rng = np.random.default_rng(1)
    n, h, w_, c = 4, 48, 80, 320
    x = rng.standard_normal((n, h, w_, c)).astype(np.float16)
    wt = (rng.standard_normal((c, c, 3, 3)) / np.sqrt(c * 9)).astype(np.float32)
    b = np.zeros(c, np.float32)
    _, ms = engine.op_conv2d(x, wt, b, 3, 1, 1, act=True, iters=20)
    flops = 2.0 * n * h * w_ * c * c * 9
    print('\nconv 3x3 320->320 @48x80 bs{}: {:.3f} ms, {:.1f} TFLOP/s'.format(n, ms, flops / ms / 1e9))
    assert ms > 0
