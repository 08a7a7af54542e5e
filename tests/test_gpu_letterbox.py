"""
GPU parity of K1 (fused letterbox + pad + /255 + space-to-depth) through the C ABI (`mdb_op_letterbox`):
bit-exact against the oracle, which itself is pinned to the real cv2.resize (tests/test_oracle_letterbox.py).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(1080, 1920), (450, 600), (450, 800), (600, 800), (768, 1280), (2160, 3840), (1281, 1919), (333, 77),
          (64, 64), (3000, 4000)]


@pytest.fixture(scope='module')
def engine():
    from megadetector_b200 import _capi
    e = _capi.Engine(_capi.make_config(max_batch=2))
    yield e
    e.close()


@pytest.mark.parametrize('hw', SHAPES, ids=['%dx%d' % s for s in SHAPES])
def test_letterbox_bit_exact(engine, hw):
    from oracle import letterbox as olb
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    want_img, _, _ = olb.letterbox(img)
    want = olb.to_stem_layout_fp16(want_img)
    got = engine.op_letterbox(img)
    assert got.shape == want.shape
    neq = got.view(np.uint16) != want.view(np.uint16)
    assert not neq.any(), '{} of {} fp16 values differ; first at {}'.format(int(neq.sum()), neq.size, np.argwhere(neq)[0].tolist())


def test_letterbox_strided_rows(engine):
    from oracle import letterbox as olb
    rng = np.random.default_rng(5)
    big = rng.integers(0, 256, size=(500, 700, 3), dtype=np.uint8)
    img = big[10:460, 20:620]  # row stride 2100 bytes, width 600
    want = olb.to_stem_layout_fp16(olb.letterbox(np.ascontiguousarray(img))[0])
    assert img.strides[0] == 2100
    from megadetector_b200 import _capi
    import ctypes as C
    g = _capi.letterbox_geometry(engine.cfg, 450, 600)
    out = np.empty((g['out_hw'][0] // 2, g['out_hw'][1] // 2, 48), np.float16)
    hw = (C.c_int32 * 2)()
    rc = engine.lib.mdb_op_letterbox(engine.h, img.ctypes.data, 450, 600, img.strides[0], out.ctypes.data, hw)
    assert rc == 0
    assert np.array_equal(out.view(np.uint16), want.view(np.uint16))
