#!/bin/bash
# Round-2 pass H: memcheck over the kernels written / rewritten this round, file-level and video throughput
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os, io
sys.path.insert(0, os.getcwd())
import numpy as np
from megadetector_b200.detector import B200Detector
from megadetector_b200.synth import synth_image
from PIL import Image
imgs = [synth_image(i, 270, 480) for i in range(3)] + [synth_image(3, 300, 400)]
for mode in ('classic', 'modern'):
    det = B200Detector('synthetic://tiny?seed=0&profile=contractive', {'max_batch': 4, 'compatibility_mode': mode})
    r = det.generate_detections_one_batch(imgs, ['a', 'b', 'c', 'd'], detection_threshold=1e-5)          # K1 (both modes), dense NMS rounds
    r2 = det.generate_detections_one_batch(imgs[:2], ['a', 'b'], detection_threshold=0.005, augment=True)  # TTA kernels
    b = io.BytesIO(); Image.fromarray(imgs[0]).save(b, format='JPEG', quality=90)
    r3 = det.generate_detections_from_jpeg_files([('j.jpg', b.getvalue())])                                # nvJPEG path + rotate
    print(mode, len(r), len(r2), len(r3), r3[0]['detections'] is not None)
    det.close()
e = __import__('megadetector_b200._capi', fromlist=['x'])
eng = e.Engine(e.make_config(max_batch=2))
rng = np.random.default_rng(0)
p = rng.random((2, 20000, 8), dtype=np.float32); p[:, :, 4] = (rng.integers(1, 3, size=(2, 20000)) / 3.0)
print('nms', eng.op_nms(p, 1e-5, 0.45)[1])
print('up', eng.op_upsample2x(rng.standard_normal((2, 5, 7, 16)).astype(np.float16)).shape, 'sppf', eng.op_sppf_pool(rng.standard_normal((2, 12, 20, 64)).astype(np.float16)).shape)
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"
grep -v "^Loading\|^No image\|^Using model\|^PTDetector" gpurun_out/r02_sanitizer_memcheck.log | tail -12
timeout 900 python tools/bench_files_video.py 2> gpurun_out/r02_files_video.err | tee gpurun_out/r02_files_video.jsonl | cut -c1-400
tail -3 gpurun_out/r02_files_video.err
