"""Runs on the GPU box: B200Detector records for three sample-sized synthetic images (the shapes of the reference's
images/nacti.jpg, idaho-camera-traps.jpg, orinoquia-thumb-web.jpg: BASELINE config 1) -> gpurun_out/b200_results_sample.json.
The committed copy (tests/golden/) is validated on the CPU box with the REFERENCE's own write_results_to_file and
validate_batch_results (tests/test_reference_schema.py)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megadetector_b200.detector import B200Detector
from megadetector_b200.synth import synth_image

det = B200Detector('synthetic://yolov5x6?seed=0', {'max_batch': 4})
shapes = [(450, 600), (450, 800), (600, 800)]
names = ['nacti.jpg', 'idaho-camera-traps.jpg', 'orinoquia-thumb-web.jpg']
imgs = [synth_image(10 + i, h, w) for i, (h, w) in enumerate(shapes)]
res = det.generate_detections_one_batch(imgs, names, detection_threshold=1e-5)
# what run_detector_batch._process_batch does next: filter on the (truncated) confidence (run_detector_batch.py:766-767)
for r in res:
    r['detections'] = [d for d in r['detections'] if d['conf'] >= 0.005]
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/b200_results_sample.json', 'w'), indent=1)
print('wrote', sum(len(r['detections']) for r in res), 'detections')
