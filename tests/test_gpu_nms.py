"""
GPU parity of K6-K8 (threshold + per-class NMS + top-k) through the C ABI (`mdb_op_nms`):
bit-exact against (a) fixtures produced by the reference's own nms() (tests/golden/nms_cases.npz) and
(b) the numpy oracle on larger seeded inputs, including kept anchor indices.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import assert_anchors_equal_modulo_ties, assert_dets_equal_modulo_ties

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def engine():
    from megadetector_b200 import _capi
    e = _capi.Engine(_capi.make_config(max_batch=4))
    yield e
    e.close()


def test_nms_matches_reference_goldens(engine):
    z = np.load(os.path.join(GOLDEN, 'nms_cases.npz'))
    names = sorted({k.split('/')[0] for k in z.files})
    for name in names:
        pred = z[name + '/pred']
        thr, iou = z[name + '/thr']
        det, cnt, anc = engine.op_nms(pred, thr, iou)
        for b in range(pred.shape[0]):
            gold = z['{}/det{}'.format(name, b)]
            assert cnt[b] == gold.shape[0], '{}[{}]: {} detections, reference has {}'.format(name, b, cnt[b], gold.shape[0])
            assert_dets_equal_modulo_ties(gold, det[b, :cnt[b]], '{}[{}]'.format(name, b))


def _random_pred(rng, n, ncl):
    centers = rng.uniform(50, 1200, size=(ncl, 2)).astype(np.float32)
    sizes = rng.uniform(20, 400, size=(ncl, 2)).astype(np.float32)
    k = rng.integers(0, ncl, size=n)
    xy = centers[k] + rng.normal(0, 12, size=(n, 2)).astype(np.float32)
    wh = np.abs(sizes[k] * rng.uniform(0.7, 1.3, size=(n, 2)).astype(np.float32))
    obj = (rng.uniform(0, 1, size=n) ** 8).astype(np.float32)
    cls = rng.dirichlet([0.6, 0.6, 0.6], size=n).astype(np.float32)
    return np.concatenate([xy, wh, obj[:, None], cls], 1).astype(np.float32)


@pytest.mark.parametrize('anchors,ncl,thr', [(61200, 40, 1e-5), (20000, 400, 0.005), (5000, 3, 1e-5), (1, 1, 1e-5)])
def test_nms_matches_oracle_at_full_size(engine, anchors, ncl, thr):
    from oracle import nms as onms
    rng = np.random.default_rng(anchors + ncl)
    pred = np.stack([_random_pred(rng, anchors, ncl) for _ in range(2)])
    det, cnt, anc = engine.op_nms(pred, thr, 0.45)
    for b in range(2):
        odet, oanc = onms.nms_one_image(pred[b], thr, 0.45)
        assert cnt[b] == odet.shape[0]
        assert_dets_equal_modulo_ties(odet, det[b, :cnt[b]], 'img{}'.format(b))
        # kept anchor indices: identical as sets inside tie groups -> compare sorted per confidence value
        assert_anchors_equal_modulo_ties(odet[:, 4], oanc, det[b, :cnt[b], 4], anc[b, :cnt[b]], 'img{}'.format(b))


def test_nms_empty_and_all_below_threshold(engine):
    pred = np.zeros((2, 256, 8), np.float32)
    det, cnt, anc = engine.op_nms(pred, 1e-5, 0.45)
    assert cnt.tolist() == [0, 0]


@pytest.mark.parametrize('n_cand,levels', [(1, 40), (31, 40), (511, 40), (512, 40), (513, 40), (1025, 40), (8191, 40),
                                           (8192, 40), (8193, 40), (30000, 40), (30000, 4000), (20000, 2), (9000, 1), (36000, 1), (36000, 3)])
def test_nms_candidate_count_boundaries(engine, n_cand, levels):
    """Window (512) and shared-memory round (8192) boundaries of the kernel, with many exactly equal confidences (ties must
    resolve to the lower anchor index, torchvision's stable order), a batch whose images have different list lengths, long
    lists cut into confidence rounds, and tie groups longer than a round (global-memory fallback sort)."""
    from oracle import nms as onms
    rng = np.random.default_rng(n_cand + levels)
    anchors = 40000
    preds = []
    for b, nc_b in enumerate((n_cand, max(1, n_cand // 3), 0)):
        p = _random_pred(rng, anchors, 60)
        obj = np.full(anchors, 1e-7, np.float32)
        sel = rng.permutation(anchors)[:nc_b]
        # quantised confidences: large tie groups
        obj[sel] = (rng.integers(1, levels + 1, size=nc_b) / float(levels + 1)).astype(np.float32)
        p[:, 4] = obj
        p[:, 5:] = np.float32(0.5) * (1 + np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=anchors)])
        preds.append(p)
    pred = np.stack(preds)
    det, cnt, anc = engine.op_nms(pred, 1e-5, 0.45)
    for b in range(pred.shape[0]):
        odet, oanc = onms.nms_torch_indexed(pred[b], 1e-5, 0.45)
        assert cnt[b] == odet.shape[0], (b, cnt[b], odet.shape[0])
        assert_dets_equal_modulo_ties(odet, det[b, :cnt[b]], 'img{}'.format(b))
        assert_anchors_equal_modulo_ties(odet[:, 4], oanc, det[b, :cnt[b], 4], anc[b, :cnt[b]], 'img{}'.format(b))
