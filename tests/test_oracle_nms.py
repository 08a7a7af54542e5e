"""CPU: the numpy NMS oracle must reproduce fixtures produced by the reference's own nms()."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import assert_dets_equal_modulo_ties
from oracle import nms as onms


def _cases():
    z = np.load(os.path.join(GOLDEN, 'nms_cases.npz'))
    names = sorted({k.split('/')[0] for k in z.files})
    return z, names


def test_oracle_matches_reference_goldens():
    z, names = _cases()
    assert 'synthetic' in names and len(names) >= 6
    total = 0
    for name in names:
        pred = z[name + '/pred']
        thr, iou = z[name + '/thr']
        res = onms.nms(pred, conf_thres=thr, iou_thres=iou)
        for b, (det, anc) in enumerate(res):
            gold = z['{}/det{}'.format(name, b)]
            assert_dets_equal_modulo_ties(gold, det, '{}[{}]'.format(name, b))
            total += det.shape[0]
            # anchor indices point at rows that reproduce the box
            if det.shape[0]:
                rows = pred[b][anc]
                assert np.array_equal(rows[:, 0] - rows[:, 2] / np.float32(2), det[:, 0])
    assert total > 1000


def test_known_answers_of_reference_synthetic_test():
    # megadetector/tests/test_nms_synthetic.py:136-335 expectations
    z, _ = _cases()
    det, _ = onms.nms(z['synthetic/pred'], conf_thres=0.3, iou_thres=0.5)[0]
    assert det.shape[0] == 8
    confs = np.round(det[:, 4].astype(np.float64), 3).tolist()
    assert confs == [0.855, 0.81, 0.72, 0.63, 0.63, 0.63, 0.54, 0.51]
    assert sorted(det[:, 5].astype(int).tolist()) == [0, 0, 0, 0, 0, 0, 0, 1]


def test_threshold_is_compared_in_float32():
    # torch: (float32 tensor > python float) casts the scalar to float32 (measured, torch 2.11)
    x = np.zeros((1, 2, 8), np.float32)
    x[0, :, :4] = [[10, 10, 4, 4], [50, 50, 4, 4]]
    x[0, :, 4] = np.float32(0.3)  # == float32(0.3) > 0.3 in double, but not in float32
    x[0, :, 5] = 1.0
    det, _ = onms.nms(x, conf_thres=0.3)[0]
    assert det.shape[0] == 0


def test_torchvision_backed_variant_agrees():
    import torch
    z, names = _cases()
    for name in names:
        pred = z[name + '/pred']
        thr, iou = z[name + '/thr']
        res = onms.nms_torch(torch.from_numpy(pred.copy()), conf_thres=thr, iou_thres=iou)
        for b, det in enumerate(res):
            assert_dets_equal_modulo_ties(z['{}/det{}'.format(name, b)], det.numpy(), '{}[{}]'.format(name, b))


def test_indexed_torch_variant_agrees_with_goldens_and_numpy_anchors():
    from helpers import assert_anchors_equal_modulo_ties
    z, names = _cases()
    for name in names:
        pred = z[name + '/pred']
        thr, iou = z[name + '/thr']
        for b in range(pred.shape[0]):
            det, anc = onms.nms_torch_indexed(pred[b], float(thr), float(iou))
            assert_dets_equal_modulo_ties(z['{}/det{}'.format(name, b)], det, '{}[{}]'.format(name, b))
            ndet, nanc = onms.nms_one_image(pred[b], thr, iou)
            assert_anchors_equal_modulo_ties(ndet[:, 4], nanc, det[:, 4], anc, '{}[{}]'.format(name, b))
