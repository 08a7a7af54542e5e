"""Shared comparison helpers for the parity tests."""
import numpy as np


def assert_dets_equal_modulo_ties(det_a, det_b, what='', max_det=300):
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
            if i == len(conf) and len(conf) >= max_det:
                # a tie group cut by max_det: which of the equal-confidence boxes survive is implementation-defined
                # in the reference (unstable argsort then [:max_det], pytorch_detector.py:597-602)
                break
            ga, gb = det_a[start:i], det_b[start:i]
            assert canon(ga) == canon(gb), '{}: rows differ in tie group at {}..{}'.format(what, start, i)
            start = i


def assert_anchors_equal_modulo_ties(conf_a, anc_a, conf_b, anc_b, what='', max_det=300):
    """Kept anchor indices must agree as sets per confidence value (a tie group cut by max_det is exempt)."""
    a = sorted(zip(conf_a.tolist(), anc_a.tolist()))
    b = sorted(zip(conf_b.tolist(), anc_b.tolist()))
    if len(a) >= max_det and len(a) == len(b):
        cut = min(conf_a.min(), conf_b.min())
        a = [t for t in a if t[0] != cut]
        b = [t for t in b if t[0] != cut]
    assert a == b, '{}: kept anchor indices differ'.format(what)
