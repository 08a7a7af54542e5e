"""CPU: the letterbox oracle is pinned bit-for-bit to the real cv2 (the library the reference executes)."""
import numpy as np
import pytest

from oracle import letterbox as olb

cv2 = pytest.importorskip('cv2')


@pytest.mark.parametrize('shape', [(1080, 1920, 720, 1280), (450, 600, 960, 1280), (450, 800, 720, 1280),
                                   (37, 53, 100, 91), (2160, 3840, 720, 1280), (480, 640, 960, 1280),
                                   (1440, 2560, 720, 1280), (201, 300, 100, 150), (720, 1280, 722, 1283)])
def test_resize_matches_cv2(shape):
    h, w, dh, dw = shape
    rng = np.random.default_rng(h * 31 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(ref, olb.resize_linear_u8(img, dw, dh))


def _yolov5_letterbox(im, new_shape=(1280, 1280), color=(114, 114, 114), auto=True, scaleup=True, stride=64):
    # the published yolov5 utils/augmentations.py::letterbox, built on the real cv2 (3P, restated for the pin)
    shape = im.shape[:2]
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    if shape[::-1] != new_unpad:
        im = cv2.resize(im, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return cv2.copyMakeBorder(im, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color), (r, r), (dw, dh)


@pytest.mark.parametrize('hw', [(1080, 1920), (450, 600), (450, 800), (600, 800), (1281, 1919), (333, 77), (64, 64),
                                (1280, 1280), (1280, 960)])
def test_letterbox_matches_cv2_pipeline(hw):
    rng = np.random.default_rng(hw[0] + hw[1])
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    want, ratio, pad = _yolov5_letterbox(img)
    got, r2, p2 = olb.letterbox(img)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert ratio == r2 and tuple(float(x) for x in pad) == tuple(float(x) for x in p2)
    assert got.shape[0] % 64 == 0 and got.shape[1] % 64 == 0


def test_benchmark_shape_geometry():
    # SURVEY.md par. 8a: 1920x1080 -> new_unpad (1280, 720), pad (0, 24) -> 768 x 1280
    g = olb.letterbox_params(1080, 1920)
    assert g['new_unpad'] == (1280, 720) and g['out_hw'] == (768, 1280) and g['top'] == 24 and g['bottom'] == 24
