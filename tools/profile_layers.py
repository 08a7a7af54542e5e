"""Per-layer device times of the benchmark workload (runs on the GPU box): python tools/profile_layers.py [batch] [out.csv]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from megadetector_b200.detector import B200Detector
from megadetector_b200.synth import synth_image

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out = sys.argv[2] if len(sys.argv) > 2 else None
det = B200Detector('synthetic://yolov5x6?seed=0', {'max_batch': B})
det.engine.set_pipeline_split(1)
imgs = [synth_image(i % 8) for i in range(B)]
det.engine.infer_batch(imgs, 1e-5, want_anchors=False)
rows = det.engine.profile_steps(5)
tot_ms = sum(r[1] for r in rows)
tot_macs = sum(r[2] for r in rows)
lines = ['step,ms,GMAC,TFLOPs,share']
for name, ms, macs in rows:
    lines.append('{},{:.4f},{:.3f},{:.1f},{:.4f}'.format(name, ms, macs / 1e9, 2 * macs / ms / 1e9 if ms > 0 else 0, ms / tot_ms))
lines.append('TOTAL,{:.3f},{:.1f},{:.1f},1.0'.format(tot_ms, tot_macs / 1e9, 2 * tot_macs / tot_ms / 1e9))
txt = '\n'.join(lines)
print(txt)
print('images/s (network only): {:.1f}'.format(B / tot_ms * 1000))
if out:
    open(out, 'w').write(txt + '\n')
