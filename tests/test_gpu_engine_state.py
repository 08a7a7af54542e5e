"""
Engine state that the reference does not have and that must therefore be invisible:
  * the per-shape plan cache is bounded (LRU): a data set with many letterboxed shapes must not exhaust HBM, and an evicted
    shape must give the same records when it comes back (pytorch_detector.py:1226-1233 groups by shape on every call);
  * two engines on two GPUs in one process are independent (`detector_options['device']`, pytorch_detector.py:863-864;
    "one handle per GPU ... different handles are independent", include/mdb200.h).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(270, 480), (480, 270), (300, 400), (400, 300), (256, 512), (512, 256), (360, 360)]


def _records(det, imgs):
    return det.generate_detections_one_batch(imgs, ['i%d' % i for i in range(len(imgs))], detection_threshold=0.001)


def test_plan_cache_eviction_is_invisible():
    from megadetector_b200.detector import B200Detector
    from megadetector_b200.synth import synth_image
    imgs = [synth_image(i, *SHAPES[i % len(SHAPES)]) for i in range(2 * len(SHAPES))]
    ref_det = B200Detector('synthetic://tiny?seed=0', {'max_batch': 4})
    ref = _records(ref_det, imgs)
    ref_det.close()
    old = os.environ.get('MDB_MAX_PLANS')
    os.environ['MDB_MAX_PLANS'] = '2'
    try:
        det = B200Detector('synthetic://tiny?seed=0', {'max_batch': 4})
    finally:
        if old is None:
            del os.environ['MDB_MAX_PLANS']
        else:
            os.environ['MDB_MAX_PLANS'] = old
    # all shapes in one call (7 letterboxed shapes through a cache of 2), then one shape at a time, then again
    assert _records(det, imgs) == ref
    for k in range(len(SHAPES)):
        sub = [i for i in range(len(imgs)) if i % len(SHAPES) == k]
        got = det.generate_detections_one_batch([imgs[i] for i in sub], ['i%d' % i for i in sub], detection_threshold=0.001)
        assert got == [ref[i] for i in sub]
    assert _records(det, imgs) == ref
    det.close()


def test_two_engines_on_two_devices_in_one_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    from megadetector_b200.detector import B200Detector
    from megadetector_b200.synth import synth_image
    imgs = [synth_image(i, 1080, 1920) for i in range(4)]
    d0 = B200Detector('synthetic://tiny?seed=0', {'max_batch': 4, 'device': 'cuda:0'})
    d1 = B200Detector('synthetic://tiny?seed=0', {'max_batch': 4, 'device': 'cuda:1'})
    r0 = _records(d0, imgs)
    r1 = _records(d1, imgs)  # the conv kernel needs > 48 KB of dynamic shared memory on THIS device too
    r0b = _records(d0, imgs)
    assert all(r['detections'] is not None for r in r1), r1[0]
    assert r0 == r1 == r0b
    d0.close()
    d1.close()
