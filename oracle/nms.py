"""
ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the reference's confidence filter + per-class NMS:

  * `nms(prediction, conf_thres, iou_thres=0.45, max_det=300)`
        /root/reference/megadetector/detection/pytorch_detector.py:502-610
  * `torchvision.ops.nms` (CPU kernel semantics, torchvision 0.26): stable sort by score descending,
    suppress when float32 `inter / (area_i + area_j - inter)` promoted to double is `> iou_threshold`.

Pinning (tests/test_oracle_nms.py): the reference's own `nms()` imported from /root/reference is run here
on the known-answer vectors of megadetector/tests/test_nms_synthetic.py:59-131 and on seeded random
predictions; the resulting fixtures are committed under tests/golden/ (generator:
tools/make_golden_nms.py) and this restatement must reproduce them bit-for-bit, including the kept
anchor indices.

Semantics worth restating because the GPU kernels must match them:
  * `x > conf_thres` on a float32 tensor compares against float32(conf_thres)  (measured with torch 2.11)
  * class confidence = float32(p_c * obj); best class = first maximum (lowest class index on ties)
  * xywh -> xyxy: `xc - w / 2.0` in float32
  * classes processed in ascending order; within a class boxes keep their anchor order, and the sort is
    stable, so "equal score -> lower anchor index first"
  * the final cross-class `argsort(descending=True)` is NOT stable in torch; ties among kept confidences
    are therefore order-undefined in the reference and compared as sets by the parity tests
"""

import numpy as np


def _nms_single_class(boxes, scores, iou_thres):
    """torchvision.ops.nms CPU semantics; returns kept indices in descending-score order."""
    n = boxes.shape[0]
    order = np.argsort(-scores.astype(np.float64), kind='stable')
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1).astype(np.float32) * (y2 - y1).astype(np.float32)
    suppressed = np.zeros(n, bool)
    keep = []
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        rest = rest[~suppressed[rest]]
        if rest.size == 0:
            continue
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), (xx2 - xx1).astype(np.float32))
        h = np.maximum(np.float32(0), (yy2 - yy1).astype(np.float32))
        inter = (w * h).astype(np.float32)
        with np.errstate(divide='ignore', invalid='ignore'):
            ovr = (inter / ((areas[i] + areas[rest]).astype(np.float32) - inter).astype(np.float32)).astype(np.float32)
        suppressed[rest[ovr.astype(np.float64) > float(iou_thres)]] = True
    return np.asarray(keep, np.int64)


def nms_one_image(x, conf_thres, iou_thres=0.45, max_det=300):
    """
    x: float32 [num_anchors, 5 + nc].  Returns (det [N,6] float32 = x1,y1,x2,y2,conf,cls ; anchor_idx [N] int64),
    rows ordered by conf descending (ties: class ascending, then per-class NMS order).
    """
    x = np.asarray(x, np.float32)
    thr = np.float32(conf_thres)
    idx = np.nonzero(x[:, 4] > thr)[0]
    if idx.size == 0:
        return np.zeros((0, 6), np.float32), np.zeros((0,), np.int64)
    xs = x[idx]
    half = np.float32(2.0)
    box = np.empty((xs.shape[0], 4), np.float32)
    box[:, 0] = xs[:, 0] - xs[:, 2] / half
    box[:, 1] = xs[:, 1] - xs[:, 3] / half
    box[:, 2] = xs[:, 0] + xs[:, 2] / half
    box[:, 3] = xs[:, 1] + xs[:, 3] / half
    class_conf = (xs[:, 5:] * xs[:, 4:5]).astype(np.float32)
    best_idx = np.argmax(class_conf, axis=1)  # first maximum
    best = class_conf[np.arange(class_conf.shape[0]), best_idx]
    m = best > thr
    if m.sum() == 0:
        return np.zeros((0, 6), np.float32), np.zeros((0,), np.int64)
    box, best, best_idx, idx = box[m], best[m], best_idx[m], idx[m]
    dets, anchors = [], []
    for c in np.unique(best_idx):
        cm = best_idx == c
        cb, cs, ci = box[cm], best[cm], idx[cm]
        keep = _nms_single_class(cb, cs, iou_thres)
        if keep.size:
            d = np.concatenate([cb[keep], cs[keep, None], np.full((keep.size, 1), c, np.float32)], 1)
            dets.append(d.astype(np.float32))
            anchors.append(ci[keep])
    if not dets:
        return np.zeros((0, 6), np.float32), np.zeros((0,), np.int64)
    det = np.concatenate(dets, 0)
    anc = np.concatenate(anchors, 0)
    order = np.argsort(-det[:, 4].astype(np.float64), kind='stable')
    det, anc = det[order], anc[order]
    if det.shape[0] > max_det:
        # NB: with exact ties straddling the cut the reference's choice is implementation-defined
        det, anc = det[:max_det], anc[:max_det]
    return det, anc


def nms(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300):
    """prediction: float32 [B, num_anchors, 5+nc] -> list of (det, anchor_idx) per image."""
    return [nms_one_image(prediction[b], conf_thres, iou_thres, max_det) for b in range(prediction.shape[0])]


def nms_torch(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300):
    """
    Same algorithm with the libraries the reference itself calls (ATen ops + torchvision.ops.nms C++ kernel),
    used where the oracle is TIMED as the CPU baseline so that the baseline is not handicapped by the numpy
    loop above.  prediction: torch float32 [B, anchors, 5+nc] -> list of [N,6] tensors.  Checked against the
    numpy restatement and the reference goldens in tests/test_oracle_nms.py.
    """
    import torch
    import torchvision
    out = []
    for x in prediction:
        x = x[x[:, 4] > conf_thres]
        if x.shape[0] == 0:
            out.append(torch.zeros((0, 6)))
            continue
        half = x[:, 2:4] / 2.0
        box = torch.cat((x[:, 0:2] - half, x[:, 0:2] + half), 1)
        conf, cls = (x[:, 5:] * x[:, 4:5]).max(1)
        m = conf > conf_thres
        box, conf, cls = box[m], conf[m], cls[m]
        parts = []
        for c in cls.unique():
            cm = cls == c
            keep = torchvision.ops.nms(box[cm], conf[cm], iou_thres)
            parts.append(torch.cat((box[cm][keep], conf[cm][keep, None], torch.full((len(keep), 1), float(c))), 1))
        if not parts:
            out.append(torch.zeros((0, 6)))
            continue
        d = torch.cat(parts, 0)
        d = d[d[:, 4].argsort(descending=True)][:max_det]
        out.append(d)
    return out


def nms_torch_indexed(pred, conf_thres=0.25, iou_thres=0.45, max_det=300):
    """
    `nms_torch` for ONE image that also returns the kept anchor indices: (det [N,6] float32 numpy, anchor_idx [N] int64).
    Same ATen ops + torchvision.ops.nms C++ kernel the reference calls (pytorch_detector.py:529-606); used by the parity
    tests where every anchor is a candidate (61 200 per image) and the numpy loop above would take minutes.  Checked
    against `nms_one_image` in tests/test_oracle_nms.py.  Ties among equal confidences: torch's final argsort is
    unstable, so callers compare tie groups as sets (tests/helpers.py).
    """
    import torch
    import torchvision
    x = torch.as_tensor(pred, dtype=torch.float32)
    idx = torch.nonzero(x[:, 4] > conf_thres).flatten()
    x = x[idx]
    if x.shape[0] == 0:
        return np.zeros((0, 6), np.float32), np.zeros((0,), np.int64)
    half = x[:, 2:4] / 2.0
    box = torch.cat((x[:, 0:2] - half, x[:, 0:2] + half), 1)
    conf, cls = (x[:, 5:] * x[:, 4:5]).max(1)
    m = conf > conf_thres
    box, conf, cls, idx = box[m], conf[m], cls[m], idx[m]
    parts, anchors = [], []
    for c in cls.unique():
        cm = cls == c
        keep = torchvision.ops.nms(box[cm], conf[cm], iou_thres)
        parts.append(torch.cat((box[cm][keep], conf[cm][keep, None], torch.full((len(keep), 1), float(c))), 1))
        anchors.append(idx[cm][keep])
    if not parts:
        return np.zeros((0, 6), np.float32), np.zeros((0,), np.int64)
    d = torch.cat(parts, 0)
    a = torch.cat(anchors, 0)
    order = torch.argsort(d[:, 4].double(), descending=True, stable=True)[:max_det]
    return d[order].numpy(), a[order].numpy()
