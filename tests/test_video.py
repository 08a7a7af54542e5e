"""
Video path (megadetector_b200/video.py) against the reference's own frame loop.

CPU: `run_detector_on_video` (frames batched) must return exactly what the reference's
`video_utils.run_callback_on_frames` (/root/reference/megadetector/detection/video_utils.py:332-468, imported unmodified from
baseline/_ref) returns with the per-frame detector callback of process_video.py:158-164, for every sampling mode; the
MD-format assembly must equal process_video.py:214-258.  A stand-in detector keeps this test off the GPU.
GPU: the same with the real B200Detector on a synthetic 4K clip.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'baseline', '_ref'), os.path.join(ROOT, 'tests', 'stubs')):
    if p not in sys.path:
        sys.path.insert(0, p)

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'megadetector')),
                               reason='baseline/_ref is not installed (tools/install_reference.sh)')


class FakeDetector:
    def generate_detections_one_batch(self, imgs, ids, detection_threshold=1e-5, image_size=None, augment=False, verbose=False):
        out = []
        for im, f in zip(imgs, ids):
            c = float(int(np.asarray(im, np.int64).sum()) % 997) / 997.0
            out.append({'file': f, 'max_detection_conf': c, 'detections': [{'category': '1', 'conf': c, 'bbox': [0.1, 0.1, 0.2, 0.2]}]})
        return out

    def generate_detections_one_image(self, im, image_id='unknown', detection_threshold=1e-5, image_size=None, augment=False,
                                      verbose=False):
        return self.generate_detections_one_batch([im], [image_id], detection_threshold)[0]


def _write_video(path, n_frames, h, w, fps=30.0):
    import cv2
    from megadetector_b200.synth import synth_image
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*'mp4v'), fps, (w, h))
    assert vw.isOpened()
    base = synth_image(7, h, w)
    for i in range(n_frames):
        f = np.roll(base, 7 * i, axis=1).copy()
        f[:16, :16] = (i * 9) % 255
        vw.write(f[:, :, ::-1])
    vw.release()
    return path


@needs_ref
@pytest.mark.parametrize('every', [None, 1, 4, 10, -0.2])
def test_batched_video_loop_equals_reference_frame_loop(tmp_path, every):
    from megadetector.detection import video_utils
    from megadetector_b200 import video
    v = _write_video(str(tmp_path / 'clip.mp4'), 23, 120, 160)
    det = FakeDetector()
    ref = video_utils.run_callback_on_frames(v, lambda im, iid: det.generate_detections_one_image(im, iid), every_n_frames=every)
    got = video.run_detector_on_video(det, v, every_n_frames=every, batch_size=5)
    assert got['frame_filenames'] == ref['frame_filenames'] and got['frame_rate'] == ref['frame_rate']
    assert got['results'] == ref['results'] and len(got['results']) > 0
    got2 = video.run_detector_on_video(det, v, frames_to_process=[0, 3, 22, 40], batch_size=2)
    ref2 = video_utils.run_callback_on_frames(v, lambda im, iid: det.generate_detections_one_image(im, iid),
                                              frames_to_process=[0, 3, 22, 40])
    assert got2['frame_filenames'] == ref2['frame_filenames'] == ['frame000000.jpg', 'frame000003.jpg', 'frame000022.jpg']
    assert got2['results'] == ref2['results']


def test_md_format_assembly_and_folder_driver(tmp_path):
    from megadetector_b200 import video
    os.makedirs(str(tmp_path / 'sub'))
    _write_video(str(tmp_path / 'a.mp4'), 9, 96, 128)
    _write_video(str(tmp_path / 'sub' / 'b.mp4'), 5, 96, 128, fps=15.0)
    with open(str(tmp_path / 'broken.mp4'), 'w') as f:
        f.write('nope')
    out = video.process_videos(FakeDetector(), str(tmp_path), str(tmp_path / 'out.json'), frame_sample=2, batch_size=4)
    by = {im['file']: im for im in out['images']}
    assert set(by) == {'a.mp4', 'broken.mp4', 'sub/b.mp4'}
    assert by['a.mp4']['frames_processed'] == [0, 2, 4, 6, 8] and by['sub/b.mp4']['frame_rate'] == 15.0
    assert sorted(d['frame_number'] for d in by['a.mp4']['detections']) == [0, 2, 4, 6, 8]
    confs = [d['conf'] for d in by['a.mp4']['detections']]
    assert confs == sorted(confs, reverse=True)
    assert by['broken.mp4']['detections'] is None and by['broken.mp4']['failure'] == video.FAILURE_VIDEO_OPEN
    assert all(set(d) == {'category', 'conf', 'bbox', 'frame_number'} for d in by['a.mp4']['detections'])


@needs_ref
@pytest.mark.gpu
def test_b200_detector_on_a_4k_clip_equals_the_per_frame_reference_loop(tmp_path):
    from megadetector.detection import video_utils
    from megadetector_b200 import video
    from megadetector_b200.detector import B200Detector
    v = _write_video(str(tmp_path / 'clip4k.mp4'), 31, 2160, 3840)
    det = B200Detector('synthetic://tiny?seed=0', {'max_batch': 8})
    ref = video_utils.run_callback_on_frames(v, lambda im, iid: det.generate_detections_one_image(im, iid, detection_threshold=0.005),
                                             every_n_frames=10)
    got = video.run_detector_on_video(det, v, every_n_frames=10, batch_size=8, detection_threshold=0.005)
    det.close()
    assert got['frame_filenames'] == ref['frame_filenames'] == ['frame000000.jpg', 'frame000010.jpg', 'frame000020.jpg', 'frame000030.jpg']
    assert got['results'] == ref['results']
    assert all(r['detections'] is not None for r in got['results'])


def _process_videos_both_ways(tmp_path, engine, make_detector):
    """The reference's own `process_video.process_videos` (process_video.py:123-280) driven through the INTEGRATION.md hook
    (one frame per detector call) must write the same "images" entries as megadetector_b200.video.process_videos (frames
    batched within each video)."""
    import json
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import run_config1
    from megadetector_b200 import video
    from megadetector_b200.synth_checkpoint import write_synthetic_checkpoint
    run_config1.install_hook()
    from megadetector.detection import process_video
    folder = tmp_path / 'videos'
    os.makedirs(str(folder / 'sub'))
    _write_video(str(folder / 'a.mp4'), 12, 270, 480)
    _write_video(str(folder / 'sub' / 'b.mp4'), 7, 300, 400, fps=15.0)
    model = str(tmp_path / 'md_synth_tiny.pt')
    write_synthetic_checkpoint('synthetic://tiny?seed=0', model)
    opt = process_video.ProcessVideoOptions()
    opt.model_file = model
    opt.input_video_file = str(folder)
    opt.output_json_file = str(tmp_path / 'ref.json')
    opt.recursive = True
    opt.frame_sample = 3
    opt.json_confidence_threshold = 0.01
    opt.detector_options = {'engine': engine, 'max_batch': 8}
    process_video.process_videos(opt)
    ref = json.load(open(opt.output_json_file))
    det = make_detector(model)
    got = video.process_videos(det, str(folder), str(tmp_path / 'got.json'), frame_sample=3, batch_size=8,
                               json_confidence_threshold=0.01, recursive=True)
    if hasattr(det, 'close'):
        det.close()
    assert got['detection_categories'] == ref['detection_categories']
    key = lambda im: im['file']
    assert sorted(got['images'], key=key) == sorted(ref['images'], key=key)
    assert sum(len(im['detections']) for im in got['images']) > 0


@needs_ref
def test_reference_process_videos_vs_batched_driver_cpu_oracle(tmp_path):
    from oracle.pt_detector import OraclePTDetector
    _process_videos_both_ways(tmp_path, 'oracle', lambda m: OraclePTDetector(m))


@needs_ref
@pytest.mark.gpu
def test_reference_process_videos_vs_batched_driver_b200(tmp_path):
    from megadetector_b200.detector import B200Detector
    _process_videos_both_ways(tmp_path, 'b200', lambda m: B200Detector(m, {'max_batch': 8}))
