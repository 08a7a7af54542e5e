"""cProfile of the reference-facing call at the benchmark workload (GPU box)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megadetector_b200 import _capi
from megadetector_b200.detector import B200Detector
from megadetector_b200.synth import synth_image

B = 64
det = B200Detector('synthetic://yolov5x6?seed=0', {'max_batch': B})
pins = [_capi.PinnedArray((1080, 1920, 3)) for _ in range(B)]
for i, p in enumerate(pins):
    synth_image(i, out=p.array)
imgs = [p.array for p in pins]
ids = ['i%d' % i for i in range(B)]
for _ in range(3):
    det.generate_detections_one_batch(imgs, ids, 1e-5)
t0 = time.perf_counter()
for _ in range(4):
    det.generate_detections_one_batch(imgs, ids, 1e-5)
print('ms per call', (time.perf_counter() - t0) / 4 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(4):
    det.generate_detections_one_batch(imgs, ids, 1e-5)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
# raw engine call, no formatting
t0 = time.perf_counter()
for _ in range(4):
    det.engine.infer_batch(imgs, 1e-5, want_anchors=False)
print('engine.infer_batch ms per call', (time.perf_counter() - t0) / 4 * 1e3, det.engine.last_timing())
