"""Shared comparison helpers for the parity tests."""
import numpy as np


def assert_dets_equal_modulo_ties(det_a, det_b, what=''):
    """
    Both [N,6] float32, sorted by conf descending.  Rows must be bit-identical as a multiset, and the
    order must agree except inside groups of exactly equal confidence (the reference's final argsort is
    unstable on ties: pytorch_detector.py:597).
    """
    assert det_a.shape == det_b.shape, '{}: shape {} vs {}'.format(what, det_a.shape, det_b.shape)
    if det_a.shape[0] == 0:
        return
    assert np.array_equal(det_a[:, 4], det_b[:, 4]), '{}: confidence sequences differ'.format(what)

    def canon(d):
        keys = [tuple(r.tobytes() for r in row.reshape(-1, 1)) for row in d]
        return sorted(keys)

    conf = det_a[:, 4]
    start = 0
    for i in range(1, len(conf) + 1):
        if i == len(conf) or conf[i] != conf[start]:
            ga, gb = det_a[start:i], det_b[start:i]
            assert canon(ga) == canon(gb), '{}: rows differ in tie group at {}..{}'.format(what, start, i)
            start = i
