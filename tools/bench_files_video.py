"""File-level throughput on the GPU box (BASELINE config 5 and the JPEG row of SURVEY.md par. 8f): what the widened rows buy.

  1. 256 synthetic 1920x1080 JPEG files -> records: PIL decode on host threads + B200 engine, vs nvJPEG decode on the GPU
     (megadetector_b200.run_batch.process_files, one GPU).
  2. a synthetic 3840x2160 @ 30 fps clip: every frame / every 10th frame through megadetector_b200.video.run_detector_on_video
     (frames batched within the video) vs the reference's per-frame callback loop shape (batch of 1 per call).
Prints one JSON line per measurement; these are side figures, not the bench metric.
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from megadetector_b200 import run_batch, video  # noqa: E402
from megadetector_b200.detector import B200Detector  # noqa: E402
from megadetector_b200.synth import synth_image  # noqa: E402


def main():
    from PIL import Image
    import cv2
    tmp = tempfile.mkdtemp(prefix='mdb200_files_')
    n_files = int(os.environ.get('N_FILES', '256'))
    files = []
    for i in range(n_files):
        p = os.path.join(tmp, 'img%04d.jpg' % i)
        Image.fromarray(synth_image(i % 64)).save(p, quality=90)
        files.append(p)
    det = B200Detector('synthetic://yolov5x6?seed=0', {'max_batch': 64})
    for mode in (False, True):
        run_batch.process_files(det, files[:64], 64, 0.005, loader_threads=16, gpu_decode=mode)  # warm-up (plans, graphs, nvJPEG)
        t0 = time.time()
        res = run_batch.process_files(det, files, 64, 0.005, loader_threads=16, gpu_decode=mode)
        dt = time.time() - t0
        assert len(res) == n_files and all('detections' in r for r in res)
        print(json.dumps({'what': 'jpeg files -> records, 1 GPU, bs=64, 16 loader threads', 'decode': 'nvJPEG (GPU)' if mode else 'PIL (host)',
                          'files': n_files, 'images_per_s': n_files / dt, 'host_cores': os.cpu_count()}), flush=True)
    # ---- video ----
    clip = os.path.join(tmp, 'clip4k.mp4')
    n_frames = int(os.environ.get('N_FRAMES', '120'))
    vw = cv2.VideoWriter(clip, cv2.VideoWriter_fourcc(*'mp4v'), 30.0, (3840, 2160))
    base = cv2.resize(synth_image(7), (3840, 2160), interpolation=cv2.INTER_LINEAR)
    for i in range(n_frames):
        vw.write(np.roll(base, 16 * i, axis=1)[:, :, ::-1])
    vw.release()
    t0 = time.time()
    cap = cv2.VideoCapture(clip)
    k = 0
    while cap.read()[0]:
        k += 1
    cap.release()
    decode_fps = k / (time.time() - t0)
    for every, bs in ((10, 32), (1, 32), (1, 1)):
        video.run_detector_on_video(det, clip, every_n_frames=every, batch_size=bs, detection_threshold=0.005)  # warm-up
        t0 = time.time()
        r = video.run_detector_on_video(det, clip, every_n_frames=every, batch_size=bs, detection_threshold=0.005)
        dt = time.time() - t0
        print(json.dumps({'what': '3840x2160@30 mp4v clip, cv2/FFmpeg decode on the host + B200 engine', 'frames_in_clip': k,
                          'every_n_frames': every, 'frames_per_detector_call': bs, 'frames_detected': len(r['results']),
                          'video_frames_per_s': k / dt, 'detected_frames_per_s': len(r['results']) / dt,
                          'decode_only_frames_per_s': decode_fps}), flush=True)
    det.close()


if __name__ == '__main__':
    main()
