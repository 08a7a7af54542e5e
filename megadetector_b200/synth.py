"""
Deterministic synthetic camera-trap-like images for the benchmark configs of BASELINE.json
("1k / 100k synthetic 1920x1080 images"; generator specified in SURVEY.md par. 8d).  There is no network,
so no real image set: a low-frequency background, per-pixel noise and a handful of filled shapes.
"""

import numpy as np


def synth_image(index, h=1080, w=1920, out=None):
    """uint8 RGB HWC image number `index` (seed = 1234 + index).  `out` may be a preallocated (pinned) array."""
    import cv2
    rng = np.random.default_rng(1234 + int(index))
    gh, gw = max(2, h // 32 + 1), max(2, w // 32)
    grid = rng.integers(0, 256, size=(gh, gw, 3), dtype=np.uint8)
    img = cv2.resize(grid, (w, h), interpolation=cv2.INTER_LINEAR).astype(np.int16)
    img += rng.integers(-8, 9, size=(h, w, 3), dtype=np.int16)
    img = np.clip(img, 0, 255).astype(np.uint8)
    n_shapes = int(rng.integers(0, 7))
    for _ in range(n_shapes):
        cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
        sw, sh = int(rng.integers(40, 601)), int(rng.integers(40, 601))
        color = tuple(int(c) for c in rng.integers(0, 256, size=3))
        if rng.integers(0, 2) == 0:
            cv2.rectangle(img, (cx - sw // 2, cy - sh // 2), (cx + sw // 2, cy + sh // 2), color, thickness=-1)
        else:
            cv2.ellipse(img, (cx, cy), (sw // 2, sh // 2), 0, 0, 360, color, thickness=-1)
    if out is not None:
        np.copyto(out, img)
        return out
    return img


def synth_pool(n, h=1080, w=1920, alloc=None):
    """List of n distinct images; `alloc(shape)` may return pinned uint8 arrays."""
    imgs = []
    for i in range(n):
        out = alloc((h, w, 3)) if alloc is not None else None
        imgs.append(synth_image(i, h, w, out=out))
    return imgs
