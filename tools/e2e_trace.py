"""Timestamps of the streaming batch API at the benchmark workload (GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megadetector_b200 import _capi
from megadetector_b200.detector import B200Detector, format_detections
from megadetector_b200.synth import synth_image

B = 64
det = B200Detector('synthetic://yolov5x6?seed=0', {'max_batch': B})
if len(sys.argv) > 1:
    det.engine.set_pipeline_split(int(sys.argv[1]))
pins = [_capi.PinnedArray((1080, 1920, 3)) for _ in range(B)]
for i, p in enumerate(pins):
    synth_image(i, out=p.array)
imgs = [p.array for p in pins]
ids = ['i%d' % i for i in range(B)]
for _ in range(3):
    det.generate_detections_one_batch(imgs, ids, 1e-5)
for mode in ['consume only', 'consume + format']:
    for rep in range(2):
        t0 = time.perf_counter()
        marks = []
        for done, d, c, s in det.engine.infer_batch_stream(imgs, 1e-5):
            t1 = time.perf_counter()
            if mode != 'consume only':
                for j in done:
                    format_detections(d[j], c[j], 1080, 1920)
            marks.append((len(done), round((t1 - t0) * 1e3, 2), round((time.perf_counter() - t1) * 1e3, 2)))
        print(mode, 'total %.2f ms' % ((time.perf_counter() - t0) * 1e3), marks, flush=True)
t0 = time.perf_counter()
for _ in range(3):
    det.generate_detections_one_batch(imgs, ids, 1e-5)
print('generate_detections_one_batch %.2f ms' % ((time.perf_counter() - t0) / 3 * 1e3))
