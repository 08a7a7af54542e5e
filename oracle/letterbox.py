"""
ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's image preparation for the MDv5 hot path:

  * `PTDetector.preprocess_image`            /root/reference/megadetector/detection/pytorch_detector.py:964-1119
  * the third-party `letterbox()` it calls    (ultralytics-yolov5 == 0.1.1, pyproject.toml:70; parameters
    restated in the reference at pytorch_detector.py:434-452)
  * `cv2.resize(..., INTER_LINEAR)` on u8     (OpenCV fixed-point bilinear, 11-bit coefficients)
  * tensor prep in `_process_batch_group`     pytorch_detector.py:1283-1306 (HWC->CHW, float32, /255.0)

Pinning: `resize_linear_u8` is checked bit-for-bit against the real `cv2.resize` (the library the
reference executes) in tests/test_oracle_letterbox.py for up-scales, down-scales, exact 2x and odd sizes.
The yolov5 `letterbox` itself is absent from /root/reference (un-vendored dependency): its parameter
arithmetic follows the reference's own restatement (pytorch_detector.py:434-452).
"""

import math

import numpy as np

INTER_RESIZE_COEF_BITS = 11
INTER_RESIZE_COEF_SCALE = 1 << INTER_RESIZE_COEF_BITS
PAD_VALUE = 114


def linear_tables(ssize, dsize, vertical):
    """
    Coefficient tables of cv::resize(INTER_LINEAR) for 8-bit images: for every destination
    index the first source index and the two 11-bit fixed-point weights.  Horizontal tables
    zero the fraction at the borders; vertical tables keep it and clamp the row indices instead
    (that asymmetry is OpenCV's, and is visible in the low bit of border rows when up-scaling).
    """
    inv_scale = dsize / ssize
    scale = 1.0 / inv_scale
    idx = np.empty(dsize, np.int32)
    c0 = np.empty(dsize, np.int16)
    c1 = np.empty(dsize, np.int16)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if not vertical:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= ssize - 1:
                s, f = ssize - 1, np.float32(0)
        idx[d] = s
        c0[d] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(INTER_RESIZE_COEF_SCALE)))
        c1[d] = int(np.rint(f * np.float32(INTER_RESIZE_COEF_SCALE)))
    return idx, c0, c1


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for HxWxC uint8, bit-exact."""
    sh, sw = src.shape[:2]
    xi, xa0, xa1 = linear_tables(sw, dw, vertical=False)
    yi, yb0, yb1 = linear_tables(sh, dh, vertical=True)
    s = src.astype(np.int32)
    xi1 = np.minimum(xi + 1, sw - 1)
    rows = s[:, xi, :] * xa0[None, :, None].astype(np.int32) + s[:, xi1, :] * xa1[None, :, None].astype(np.int32)
    y0 = np.clip(yi, 0, sh - 1)
    y1 = np.clip(yi + 1, 0, sh - 1)
    b0 = yb0[:, None, None].astype(np.int32)
    b1 = yb1[:, None, None].astype(np.int32)
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox_params(h, w, new_shape=1280, stride=64, auto=True, scaleup=True):
    """
    Geometry of yolov5 `letterbox(im, new_shape, stride=stride, auto=auto, scaleFill=False, scaleup=scaleup)`
    as called at pytorch_detector.py:1104-1109 ("classic": auto=True, scaleup=True, new_shape=1280).

    Returns dict: new_unpad (w,h), pad (dw,dh) halves as floats, top/bottom/left/right ints, ratio, out (h,w).
    """
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = (int(round(w * r)), int(round(h * r)))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = dw % stride, dh % stride
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return {
        'ratio': (r, r),
        'new_unpad': new_unpad,
        'pad': (dw, dh),
        'top': top, 'bottom': bottom, 'left': left, 'right': right,
        'out_hw': (new_unpad[1] + top + bottom, new_unpad[0] + left + right),
    }


def letterbox(img, new_shape=1280, stride=64, auto=True, scaleup=True):
    """Returns (u8 HWC letterboxed image, ratio, (dw, dh)) exactly like yolov5 `letterbox`."""
    h, w = img.shape[:2]
    g = letterbox_params(h, w, new_shape, stride, auto, scaleup)
    nw, nh = g['new_unpad']
    if (w, h) != (nw, nh):
        img = resize_linear_u8(img, nw, nh)
    oh, ow = g['out_hw']
    out = np.full((oh, ow, img.shape[2]), PAD_VALUE, np.uint8)
    out[g['top']:g['top'] + nh, g['left']:g['left'] + nw] = img
    return out, g['ratio'], g['pad']


def to_model_input(img_letterboxed):
    """pytorch_detector.py:1286-1306: HWC u8 -> CHW float32 / 255.0 (true IEEE division)."""
    x = np.ascontiguousarray(img_letterboxed.transpose(2, 0, 1)).astype(np.float32)
    return x / np.float32(255.0)


def to_stem_layout_fp16(img_letterboxed):
    """
    The engine's K1 output layout, derived from the oracle image: 2x2 space-to-depth (channel (py*2+px)*3+c, 12
    values) with the stem's three column taps packed into the channel dimension:
        out[y, x, t*16 + k] = s2d[y, x + t - 1, k]   (zero outside the image, channels 12..15 of each tap zero)
                                                       -> [H/2][W/2][48] fp16
    """
    x = (img_letterboxed.astype(np.float32) / np.float32(255.0)).astype(np.float16)
    h, w, c = x.shape
    s2d = np.zeros((h // 2, w // 2 + 2, 12), np.float16)  # one zero column on each side
    for py in range(2):
        for px in range(2):
            s2d[:, 1:-1, (py * 2 + px) * 3:(py * 2 + px) * 3 + 3] = x[py::2, px::2, :]
    out = np.zeros((h // 2, w // 2, 48), np.float16)
    for t in range(3):
        out[:, :, t * 16:t * 16 + 12] = s2d[:, t:t + w // 2, :]
    return out
