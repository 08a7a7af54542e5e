"""
Opt-in check against a REAL MegaDetector checkpoint: no MDv5 file exists offline, so this is skipped unless
    MDB_REAL_CHECKPOINT=/path/to/md_v5a.0.0.pt   (and, for the reference arm, an importable `yolov5` / `ultralytics` package)
is set.  It loads the file through megadetector_b200.checkpoint (placeholder unpickling + Conv/BN folding), runs the CPU oracle
of the path on the three sample JPEGs and, on a GPU box, the B200 engine; when the reference's own PTDetector can be constructed
(yolov5 package present) its records are compared under the md_tests tolerances (conf 0.005, coords 0.002 = two rounded
pixels; boxes are bit-exact against the CPU reference only -- torch's CUDA kernels compute `coords /= gain` as a * (1 / gain)).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT = os.environ.get('MDB_REAL_CHECKPOINT', '')

pytestmark = pytest.mark.skipif(not (CKPT and os.path.isfile(CKPT)), reason='set MDB_REAL_CHECKPOINT to a real MDv5 .pt file')


def _images():
    from megadetector_b200 import run_batch
    d = os.path.join(ROOT, 'tests', 'golden', 'config1_images')
    files = sorted(os.path.join(d, f) for f in os.listdir(d))
    return files, [run_batch.load_image(f) for f in files]


def test_real_checkpoint_loads_and_oracle_runs():
    from megadetector_b200 import checkpoint
    from oracle.pt_detector import OraclePTDetector
    w, arch = checkpoint.load_model(CKPT)
    assert arch['nc'] == 3 and len(arch['ch']) == 6
    files, imgs = _images()
    recs = OraclePTDetector(CKPT).generate_detections_one_batch(imgs, files, detection_threshold=0.005)
    assert all(r['detections'] is not None for r in recs)
    # a trained detector finds the animals of the sample images with high confidence
    assert max(r['max_detection_conf'] for r in recs) > 0.5


@pytest.mark.gpu
def test_real_checkpoint_engine_vs_oracle_and_reference():
    import json
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from config1_compare import assert_equivalent
    from megadetector_b200.detector import B200Detector
    from oracle.pt_detector import OraclePTDetector
    files, imgs = _images()
    got = B200Detector(CKPT, {'max_batch': 4}).generate_detections_one_batch(imgs, files, detection_threshold=0.005)
    ref = OraclePTDetector(CKPT).generate_detections_one_batch(imgs, files, detection_threshold=0.005)
    arms = {'b200': got, 'oracle': ref}
    try:  # the reference's own detector, when its third-party model package is installed
        for p in (os.path.join(ROOT, 'baseline', '_ref'), os.path.join(ROOT, 'tests', 'stubs')):
            sys.path.insert(0, p)
        from megadetector.detection.pytorch_detector import PTDetector
        arms['reference'] = PTDetector(CKPT, {'force_cpu': True}).generate_detections_one_batch(imgs, files, detection_threshold=0.005)
    except Exception as e:
        print('reference PTDetector unavailable:', e)
    tmp = tempfile.mkdtemp()
    paths = {}
    for k, recs in arms.items():
        paths[k] = os.path.join(tmp, k + '.json')
        with open(paths[k], 'w') as f:
            json.dump({'images': recs, 'detection_categories': {'1': 'animal', '2': 'person', '3': 'vehicle'}, 'info': {}}, f)
    assert_equivalent(paths['b200'], paths['oracle'], 'real checkpoint: B200 vs CPU oracle')
    if 'reference' in paths:
        assert_equivalent(paths['oracle'], paths['reference'], 'real checkpoint: CPU oracle vs reference PTDetector')
        assert_equivalent(paths['b200'], paths['reference'], 'real checkpoint: B200 vs reference PTDetector')
