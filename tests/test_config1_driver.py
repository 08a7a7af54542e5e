"""
BASELINE.json config 1 through the reference's OWN driver stack (VERDICT r1 item 3 / SURVEY.md par. 7 step 2):
    run_detector_batch.load_and_run_detector_batch (run_detector_batch.py:1062) -> load_detector (run_detector.py:601,
    with the INTEGRATION.md hook) -> write_results_to_file (:1546) -> validate_batch_results
on the three sample JPEGs.  The unmodified reference package comes from baseline/_ref (tools/install_reference.sh).

  * CPU (here):  the oracle arm runs now; its JSON must equal the committed CPU golden and be equivalent (md_tests
                 tolerances) to the JSON the B200 engine produced on the GPU box (tests/golden/config1_b200_gpu.json).
  * GPU (-m gpu): the B200 arm runs through the plain batch path, the image queue and the preprocessing image queue
                 (run_detector_batch.py:124-200: dict inputs from producer threads); all three JSON files must be
                 identical to each other and equivalent to the committed CPU golden.
"""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

GOLDEN_CPU = os.path.join(ROOT, 'tests', 'golden', 'config1_oracle_cpu.json')
GOLDEN_GPU = os.path.join(ROOT, 'tests', 'golden', 'config1_b200_gpu.json')

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'megadetector')),
                               reason='baseline/_ref is not installed (tools/install_reference.sh)')


def _strip(path):
    d = json.load(open(path))
    d['info'].pop('detection_completion_time', None)
    return d


@needs_ref
def test_oracle_arm_through_the_reference_driver(tmp_path):
    import run_config1
    from config1_compare import assert_equivalent
    out = str(tmp_path / 'oracle.json')
    run_config1.run('oracle', out, batch_size=1)
    assert _strip(out) == _strip(GOLDEN_CPU)
    if os.path.exists(GOLDEN_GPU):
        assert_equivalent(GOLDEN_GPU, out, 'committed B200 JSON vs fresh CPU oracle JSON')


@needs_ref
@pytest.mark.gpu
def test_b200_arm_through_the_reference_driver(tmp_path):
    import run_config1
    from config1_compare import assert_equivalent
    outs = {}
    for tag, kw in (('batch', dict(batch_size=4)),
                    ('single', dict(batch_size=1)),
                    ('queue', dict(batch_size=4, image_queue=True)),
                    ('queue_preprocess', dict(batch_size=4, image_queue=True, preprocess_on_queue=True))):
        outs[tag] = str(tmp_path / (tag + '.json'))
        run_config1.run('b200', outs[tag], **kw)
    ref = _strip(outs['batch'])
    assert any(len(im['detections']) > 0 for im in ref['images'])
    for tag in ('single', 'queue'):
        assert _strip(outs[tag]) == ref, tag
    # producer threads letterbox on the host with cv2 (run_detector_batch.py:175-183); K1 is bit-exact against cv2, so the
    # records are identical too
    assert _strip(outs['queue_preprocess']) == ref
    assert_equivalent(outs['batch'], GOLDEN_CPU, 'fresh B200 JSON vs committed CPU oracle JSON')
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'config1_b200_gpu.json'), 'w') as f:
        json.dump(json.load(open(outs['batch'])), f, indent=1)
