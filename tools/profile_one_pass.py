"""One 64-image pass of the whole path in ONE pipeline chunk (GPU box): K1 on 64 images, the network, decode and NMS on 64
images -- the launch the `ncu --set full` captures of K1 / K3 / K4 / K5 / K6-K8 are taken from (bench configuration)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megadetector_b200 import _capi
from megadetector_b200.detector import B200Detector
from megadetector_b200.synth import synth_image

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
det = B200Detector(sys.argv[2] if len(sys.argv) > 2 else 'synthetic://yolov5x6?seed=0', {'max_batch': B})
det.engine.set_pipeline_split(1)
imgs = [synth_image(i) for i in range(B)]
for _ in range(2):
    det.engine.infer_batch(imgs, 1e-5, want_anchors=False)
print('timing', det.engine.last_timing())
