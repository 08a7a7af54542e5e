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


# ---------------------------------------------------------------------------------------------------------------
# "modern" compatibility mode (pytorch_detector.py:1036-1101): INTER_AREA / INTER_LINEAR pre-resize + pad
# ---------------------------------------------------------------------------------------------------------------
def _area_tab(ssize, dsize, scale):
    """cv::computeResizeAreaTab (OpenCV imgproc/src/resize.cpp): [(dst index, src index, float32 weight)] in table order."""
    import math
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = int(math.ceil(fsx1)), int(math.floor(fsx2))
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area_u8(src, dw, dh):
    """
    cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA) for uint8 HWC images when shrinking on both axes, op for op:
    integer scales on both axes -> cv::ResizeAreaFast_ (integer block sums; 2x2: (s + 2) >> 2, else cvRound(s * (1.f / area)));
    otherwise cv::ResizeArea_ (float32 taps from computeResizeAreaTab, row buffer accumulated in table order, vertical
    accumulation `sum = beta * buf` / `sum += beta * buf`, cvRound).  Pinned against the real cv2 in
    tests/test_oracle_modern.py.
    """
    sh, sw, cn = src.shape
    scale_x = 1.0 / (dw / sw)
    scale_y = 1.0 / (dh / sh)
    isx, isy = int(round(scale_x)), int(round(scale_y))
    if abs(scale_x - isx) < np.finfo(np.float64).eps and abs(scale_y - isy) < np.finfo(np.float64).eps:
        s = src[:dh * isy, :dw * isx].astype(np.int32).reshape(dh, isy, dw, isx, cn).sum(axis=(1, 3))
        if isx == 2 and isy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        v = s.astype(np.float32) * (np.float32(1.0) / np.float32(isx * isy))
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    S = src.astype(np.float32)
    buf = np.zeros((sh, dw, cn), np.float32)
    for dx, sx, a in _area_tab(sw, dw, scale_x):
        buf[:, dx, :] = buf[:, dx, :] + S[:, sx, :] * a
    out = np.zeros((dh, dw, cn), np.float32)
    seen = set()
    for dy, sy, b in _area_tab(sh, dh, scale_y):
        if dy not in seen:
            out[dy] = b * buf[sy]
            seen.add(dy)
        else:
            out[dy] = out[dy] + b * buf[sy]
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def preprocess_modern(img, image_size=1280, stride=64, use_ceil=False):
    """
    pytorch_detector.py:1036-1101 for a non-"classic" compatibility mode.  Returns a dict with the reference's keys:
    img_processed, img_original (the PRE-RESIZED image, as in the reference), scaling_shape, letterbox_ratio, letterbox_pad,
    target_shape.
    """
    import math
    scaling_shape = img.shape
    h, w = img.shape[:2]
    resize_ratio = image_size / max(h, w)
    if resize_ratio != 1:
        if use_ceil:
            tw, th = math.ceil(w * resize_ratio), math.ceil(h * resize_ratio)
        else:
            tw, th = int(w * resize_ratio), int(h * resize_ratio)
        img = resize_linear_u8(img, tw, th) if resize_ratio > 1 else resize_area_u8(img, tw, th)
    max_dimension = max(img.shape)
    normalized_shape = [img.shape[0] / max_dimension, img.shape[1] / max_dimension]
    target_shape = np.ceil(((np.array(normalized_shape) * image_size) / stride) + 0.5).astype(int) * stride
    # yolov5 letterbox(img, new_shape=target_shape, stride, auto=False, scaleFill=False, scaleup=False)
    shape = img.shape[:2]
    r = min(target_shape[0] / shape[0], target_shape[1] / shape[1])
    r = min(r, 1.0)
    new_unpad = (int(round(shape[1] * r)), int(round(shape[0] * r)))
    dw, dh = target_shape[1] - new_unpad[0], target_shape[0] - new_unpad[1]
    dw /= 2
    dh /= 2
    lb = img
    if shape[::-1] != new_unpad:
        lb = resize_linear_u8(img, new_unpad[0], new_unpad[1])
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    out = np.full((lb.shape[0] + top + bottom, lb.shape[1] + left + right, 3), PAD_VALUE, np.uint8)
    out[top:top + lb.shape[0], left:left + lb.shape[1]] = lb
    return {'img_processed': out, 'img_original': img, 'scaling_shape': scaling_shape, 'letterbox_ratio': (r, r),
            'letterbox_pad': (float(dw), float(dh)), 'target_shape': target_shape}
