"""
CPU, in the build container only: records produced by B200Detector on a B200 (fixture
tests/golden/b200_results_sample.json, generator tools/make_sample_results.py, shapes of BASELINE config 1) are fed to
the REFERENCE's own `write_results_to_file` (run_detector_batch.py:1546) and `validate_batch_results`
(postprocessing/validate_batch_results.py:71): the drop-in's records must be a valid MegaDetector results file.
Skipped where /root/reference does not exist (the GPU box).
"""
import json
import os
import sys

import pytest

from conftest import GOLDEN, ROOT

REF = '/root/reference'
FIXTURE = os.path.join(GOLDEN, 'b200_results_sample.json')


@pytest.mark.skipif(not os.path.isdir(REF) or not os.path.exists(FIXTURE), reason='needs the reference mount and the fixture')
def test_reference_writer_and_validator_accept_our_records(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'stubs'))
    sys.path.insert(0, REF)
    try:
        from megadetector.detection.run_detector_batch import write_results_to_file
        from megadetector.postprocessing.validate_batch_results import validate_batch_results
        results = json.load(open(FIXTURE))
        assert len(results) == 3 and all('detections' in r for r in results)
        out = str(tmp_path / 'out.json')
        d = write_results_to_file(results, out, detector_file='md_v5a.0.1.pt')
        assert d['info']['format_version'] and set(d['detection_categories']) == {'1', '2', '3'}
        v = validate_batch_results(out)
        assert v['validation_results']['errors'] == [], v['validation_results']['errors']
        written = json.load(open(out))
        assert [im['file'] for im in written['images']] == sorted(r['file'] for r in results)
        for im in written['images']:
            confs = [det['conf'] for det in im['detections']]
            assert confs == sorted(confs, reverse=True)  # written in descending confidence (:1631-1637)
            assert 'max_detection_conf' not in im
            for det in im['detections']:
                assert isinstance(det['conf'], float) and det['category'] in ('1', '2', '3') and len(det['bbox']) == 4
    finally:
        sys.path.remove(REF)
        sys.path.remove(os.path.join(ROOT, 'tests', 'stubs'))
