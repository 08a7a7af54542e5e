"""e2e through B200Detector with PAGEABLE numpy images (what PIL / cv2 callers hand in) vs page-locked ones (GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from megadetector_b200 import _capi
from megadetector_b200.detector import B200Detector
from megadetector_b200.synth import synth_image

B = 64
det = B200Detector('synthetic://yolov5x6?seed=0', {'max_batch': B})
pins = [_capi.PinnedArray((1080, 1920, 3)) for _ in range(B)]
for i, p in enumerate(pins):
    synth_image(i, out=p.array)
pinned = [p.array for p in pins]
pageable = [np.array(a) for a in pinned]
ids = ['i%d' % i for i in range(B)]
for name, imgs in (('pinned', pinned), ('pageable', pageable), ('pinned', pinned), ('pageable', pageable)):
    for _ in range(2):
        det.generate_detections_one_batch(imgs, ids, 1e-5)
    t0 = time.perf_counter()
    for _ in range(4):
        det.generate_detections_one_batch(imgs, ids, 1e-5)
    dt = (time.perf_counter() - t0) / 4
    print('%-9s %.2f ms per 64 images = %.0f images/s' % (name, dt * 1e3, B / dt), flush=True)
