"""
Generates tests/golden/nms_cases.npz by running the REFERENCE's own `nms()`
(/root/reference/megadetector/detection/pytorch_detector.py:502-610) in this container.

    PYTHONPATH=/root/reference:tests/stubs python tools/make_golden_nms.py

The reference cannot travel to the GPU box, the fixtures do.  Cases: the known-answer vectors of
megadetector/tests/test_nms_synthetic.py plus seeded random predictions shaped like YOLOv5 output rows
[xc, yc, w, h, obj, p0, p1, p2] (clustered boxes, heavy-tailed objectness, duplicated rows, zero-area boxes).
"""
import os
import sys

import numpy as np
import torch

from megadetector.detection.pytorch_detector import nms as reference_nms
from megadetector.tests.test_nms_synthetic import create_synthetic_predictions


def random_prediction(rng, n, n_clusters, thr_frac):
    centers = rng.uniform(50, 1200, size=(n_clusters, 2)).astype(np.float32)
    sizes = rng.uniform(20, 400, size=(n_clusters, 2)).astype(np.float32)
    k = rng.integers(0, n_clusters, size=n)
    xy = centers[k] + rng.normal(0, 12, size=(n, 2)).astype(np.float32)
    wh = np.abs(sizes[k] * rng.uniform(0.7, 1.3, size=(n, 2)).astype(np.float32))
    obj = (rng.uniform(0, 1, size=n) ** 6).astype(np.float32)
    obj[rng.uniform(size=n) < thr_frac] *= np.float32(1e-7)
    cls = rng.dirichlet([0.6, 0.6, 0.6], size=n).astype(np.float32)
    x = np.concatenate([xy, wh, obj[:, None], cls], 1).astype(np.float32)
    # exact duplicates (identical score ties inside a class) and zero-area boxes
    dup = rng.integers(0, n, size=max(1, n // 50))
    x[rng.integers(0, n, size=dup.size)] = x[dup]
    x[rng.integers(0, n, size=max(1, n // 100)), 2:4] = 0
    return x


def main():
    out = {}
    cases = []
    pred = create_synthetic_predictions().numpy()
    cases.append(('synthetic', pred, 0.3, 0.5))
    rng = np.random.default_rng(20260924)
    for i, (n, ncl, thr, iou) in enumerate([(300, 6, 1e-5, 0.45), (1500, 25, 1e-5, 0.45), (2000, 10, 0.005, 0.45),
                                            (1200, 4, 0.3, 0.6), (900, 200, 1e-5, 0.45), (64, 2, 0.999999, 0.45)]):
        x = np.stack([random_prediction(rng, n, ncl, 0.3) for _ in range(2)])
        cases.append(('random{}'.format(i), x, thr, iou))
    for name, pred, thr, iou in cases:
        res = reference_nms(torch.from_numpy(pred.copy()), conf_thres=thr, iou_thres=iou)
        out[name + '/pred'] = pred
        out[name + '/thr'] = np.array([thr, iou], np.float64)
        for b, det in enumerate(res):
            out['{}/det{}'.format(name, b)] = det.numpy().astype(np.float32)
        print(name, pred.shape, [int(d.shape[0]) for d in res])
    dst = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'nms_cases.npz')
    np.savez_compressed(dst, **out)
    print('wrote', os.path.abspath(dst), os.path.getsize(dst), 'bytes')


if __name__ == '__main__':
    sys.exit(main())
