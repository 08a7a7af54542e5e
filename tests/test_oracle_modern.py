"""CPU: the oracle's "modern"-mode preprocessing (oracle/letterbox.py) is pinned to the real cv2 -- the library the reference
executes at /root/reference/megadetector/detection/pytorch_detector.py:1036-1062 (INTER_AREA / INTER_LINEAR pre-resize) and in
yolov5's letterbox (copyMakeBorder) -- and to a literal transcription of the reference's size arithmetic."""
import math

import cv2
import numpy as np
import pytest

from oracle import letterbox as olb

SHAPES = [(1080, 1920), (2160, 3840), (1440, 2560), (600, 800), (450, 600), (3000, 4000), (1000, 1500), (333, 517),
          (1280, 1280), (1280, 720), (100, 300), (2448, 3264), (1944, 2592)]


@pytest.mark.parametrize('hw', [s for s in SHAPES if max(s) > 1280])
def test_area_resize_matches_cv2(hw):
    rng = np.random.default_rng(hw[0] + hw[1])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    r = 1280 / max(hw)
    tw, th = int(hw[1] * r), int(hw[0] * r)
    assert np.array_equal(olb.resize_area_u8(img, tw, th), cv2.resize(img, (tw, th), interpolation=cv2.INTER_AREA))


@pytest.mark.parametrize('hw', SHAPES)
@pytest.mark.parametrize('use_ceil', [False, True])
def test_modern_preprocessing_matches_reference_arithmetic_with_cv2(hw, use_ceil):
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    image_size, stride = 1280, 64
    # --- literal transcription of pytorch_detector.py:1040-1101 with the real cv2 ---
    h, w = img.shape[:2]
    resize_ratio = image_size / max(h, w)
    ref = img
    if resize_ratio != 1:
        interp = cv2.INTER_LINEAR if resize_ratio > 1 else cv2.INTER_AREA
        if use_ceil:
            tw, th = math.ceil(w * resize_ratio), math.ceil(h * resize_ratio)
        else:
            tw, th = int(w * resize_ratio), int(h * resize_ratio)
        ref = cv2.resize(ref, (tw, th), interpolation=interp)
    max_dimension = max(ref.shape)
    normalized_shape = [ref.shape[0] / max_dimension, ref.shape[1] / max_dimension]
    target_shape = np.ceil(((np.array(normalized_shape) * image_size) / stride) + 0.5).astype(int) * stride
    dw, dh = (target_shape[1] - ref.shape[1]) / 2, (target_shape[0] - ref.shape[0]) / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    ref_lb = cv2.copyMakeBorder(ref, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))
    got = olb.preprocess_modern(img, image_size, stride, use_ceil)
    assert got['img_processed'].shape == ref_lb.shape == (target_shape[0], target_shape[1], 3)
    assert np.array_equal(got['img_processed'], ref_lb)
    assert got['letterbox_pad'] == (dw, dh) and got['letterbox_ratio'] == (1.0, 1.0)
    assert np.array_equal(got['img_original'], ref)
    # the engine's host-side geometry (C ABI) agrees
    from megadetector_b200 import _capi
    g = _capi.letterbox_geometry(_capi.make_config(compat_mode=1, use_ceil_for_resize=use_ceil), h, w)
    assert g['new_unpad'] == (ref.shape[1], ref.shape[0]) and g['out_hw'] == tuple(target_shape)
    assert (g['top'], g['bottom'], g['left'], g['right']) == (top, bottom, left, right) and g['pad'] == (dw, dh)
