"""
GPU path-equivalence tests in the style of the reference's own CLI tests
(/root/reference/megadetector/utils/md_tests.py:1116-1599: batch vs single, queue/preprocessed inputs, repeated runs must give
identical records -- `output_files_are_identical`, md_tests.py:363-413).  Our engine is deterministic and an image's result
does not depend on what else is in the batch, so every comparison here is exact.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def det():
    from megadetector_b200.detector import B200Detector
    d = B200Detector('synthetic://tiny?seed=0', {'max_batch': 4})
    yield d
    d.close()


def _images():
    from megadetector_b200.synth import synth_image
    sizes = [(270, 480), (300, 400), (480, 270), (270, 480), (64, 64), (333, 77), (1080, 1920), (270, 480), (300, 400),
             (2000, 3000), (270, 480)]
    return [synth_image(100 + i, h, w) for i, (h, w) in enumerate(sizes)], ['im%02d.jpg' % i for i in range(len(sizes))]


def test_batch_larger_than_max_batch_with_mixed_shapes_equals_single_image_calls(det):
    imgs, ids = _images()
    batch = det.generate_detections_one_batch(imgs, ids, detection_threshold=1e-5)
    assert [r['file'] for r in batch] == ids
    for im, name, r in zip(imgs, ids, batch):
        one = det.generate_detections_one_image(im, name, detection_threshold=1e-5)
        assert one == r, name
    again = det.generate_detections_one_batch(imgs, ids, detection_threshold=1e-5)
    assert again == batch  # deterministic


def test_input_container_variants_give_identical_records(det):
    from PIL import Image
    imgs, ids = _images()
    base = det.generate_detections_one_batch(imgs[:3], ids[:3], detection_threshold=1e-5)
    pil = [Image.fromarray(a) for a in imgs[:3]]
    assert det.generate_detections_one_batch(pil, ids[:3], detection_threshold=1e-5) == base
    # non-contiguous view (row stride != width*3) and a Fortran-ordered copy
    big = np.zeros((300, 520, 3), np.uint8)
    big[10:280, 20:500] = imgs[0]
    view = big[10:280, 20:500]
    assert not view.flags['C_CONTIGUOUS']
    assert det.generate_detections_one_image(view, ids[0], detection_threshold=1e-5) == base[0]
    f = np.asfortranarray(imgs[1])
    assert det.generate_detections_one_image(f, ids[1], detection_threshold=1e-5) == base[1]


def test_threshold_semantics_and_ordering(det):
    imgs, ids = _images()
    lo = det.generate_detections_one_image(imgs[0], ids[0], detection_threshold=1e-5)
    hi = det.generate_detections_one_image(imgs[0], ids[0], detection_threshold=0.2)
    confs = [d['conf'] for d in lo['detections']]
    assert confs == sorted(confs)  # ascending, like `for ... in reversed(det)` (pytorch_detector.py:1384)
    assert all(d['category'] in ('1', '2', '3') and len(d['bbox']) == 4 for d in lo['detections'])
    assert all(isinstance(d['conf'], float) for d in lo['detections'])
    assert len(lo['detections']) <= 300 and len(hi['detections']) <= len(lo['detections'])
    assert all(d['conf'] >= 0.2 - 1e-3 for d in hi['detections'])
    assert lo['max_detection_conf'] == (max(confs) if confs else 0.0)
    assert det.generate_detections_one_batch([], []) == []


def test_image_size_argument_changes_the_letterbox(det):
    imgs, ids = _images()
    a = det.generate_detections_one_image(imgs[6], ids[6], detection_threshold=1e-5)
    b = det.generate_detections_one_image(imgs[6], ids[6], detection_threshold=1e-5, image_size=640)
    c = det.generate_detections_one_image(imgs[6], ids[6], detection_threshold=1e-5)
    assert a == c and a != b
