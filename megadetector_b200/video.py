"""
Video path (SURVEY.md par. 8f-2, BASELINE.json config 5): frame sampling + detection with frames BATCHED within a video.

The reference decodes a video with cv2.VideoCapture and calls the detector once PER FRAME
(/root/reference/megadetector/detection/video_utils.py:332-468 `run_callback_on_frames`,
 /root/reference/megadetector/detection/process_video.py:158-164 `frame_callback` -> `generate_detections_one_image`);
batching frames inside a video is the author's own TODO (/root/reference/TODO.md:537-539).  Here the sampled frames of a video
are collected into batches for `generate_detections_one_batch` while a reader thread keeps decoding (cv2 releases the GIL in
`read()`), and the results are assembled into exactly the structures the reference builds:

  * `run_detector_on_video`        -> the dict `run_callback_on_frames` returns ('frame_filenames', 'frame_rate', 'results')
  * `video_results_to_md_format`   -> one entry of the results file's "images" list (process_video.py:214-258:
                                      'file', 'frame_rate', 'frames_processed', detections with 'frame_number')
  * `process_videos`               -> folder / file driver writing the MD-format JSON (megadetector-output-format.md:101-133)

Sampling semantics are the reference's: `every_n_frames` None/0/1 = every frame, N > 1 = frames 0, N, 2N, ...; a negative
value is a period in seconds, converted with int(seconds * frame_rate); frame ids are 'frame%06d.jpg'.
NVDEC is not used: the Video Codec SDK headers are not part of this image, so decode stays on the host (cv2 / FFmpeg).
"""
import json
import os
import queue
import threading
import time

VIDEO_EXTENSIONS = ('.mp4', '.avi', '.mpeg', '.mpg', '.mov', '.mkv', '.flv')  # video_utils.default_video_extensions
FAILURE_VIDEO_OPEN = 'video open failure'


def frame_number_to_filename(frame_number):
    return 'frame{:06d}.jpg'.format(frame_number)  # video_utils.py:274-279


def _resolve_every_n(every_n_frames, frame_rate):
    if every_n_frames is None:
        return None
    if every_n_frames < 0:
        return int(abs(every_n_frames) * frame_rate)
    if every_n_frames == 0:
        return 1
    return int(every_n_frames)


def run_detector_on_video(detector, input_video_file, every_n_frames=None, frames_to_process=None, batch_size=32,
                          detection_threshold=0.00001, image_size=None, allow_empty_videos=False, queue_depth=None):
    """Same return value as video_utils.run_callback_on_frames(video, frame_callback, every_n_frames, ...) with the
    reference's detector callback, computed with batched detector calls."""
    import cv2
    assert os.path.isfile(input_video_file), 'File {} not found'.format(input_video_file)
    if isinstance(frames_to_process, int):
        frames_to_process = [frames_to_process]
    if (frames_to_process is not None) and (every_n_frames is not None):
        raise ValueError('frames_to_process and every_n_frames are mutually exclusive')
    vidcap = cv2.VideoCapture(input_video_file)
    ok, first = vidcap.read()
    if not ok:
        vidcap.release()
        raise OSError('Could not open video {}'.format(input_video_file))
    n_frames = int(vidcap.get(cv2.CAP_PROP_FRAME_COUNT))
    frame_rate = vidcap.get(cv2.CAP_PROP_FPS)
    every = _resolve_every_n(every_n_frames, frame_rate)
    wanted = set(frames_to_process) if frames_to_process is not None else None
    last_wanted = max(wanted) if wanted else None
    depth = queue_depth or max(2 * batch_size, 8)
    q = queue.Queue(maxsize=depth)
    stop = threading.Event()
    # Frames are colour-converted straight into page-locked buffers owned by the engine's library: the H2D copy of a batch is
    # then an asynchronous DMA at PCIe rate instead of CUDA's staged copy from pageable memory (~4 GB/s measured: 4K frames
    # were upload-bound at 168 frames/s).  A buffer returns to the pool when its batch has been detected.
    # The pool is kept on the detector between videos of the same frame size: page-locking 2.4 GB (98 4K frames) takes
    # about a second.
    pool, free_q = None, None
    if getattr(detector, 'engine', None) is not None:
        try:
            from . import _capi
            cache = detector.__dict__.setdefault('_video_frame_pools', {})
            pool = cache.get(tuple(first.shape), [])
            while len(pool) < depth + batch_size + 2:
                pool.append(_capi.PinnedArray(first.shape))
            for k in [k for k in cache if k != tuple(first.shape)]:
                del cache[k]  # one frame size at a time
            cache[tuple(first.shape)] = pool
            free_q = queue.Queue()
            for i in range(len(pool)):
                free_q.put(i)
        except Exception:
            pool, free_q = None, None

    def to_rgb(image):
        if pool is None:
            return cv2.cvtColor(image, cv2.COLOR_BGR2RGB), -1
        i = free_q.get()
        cv2.cvtColor(image, cv2.COLOR_BGR2RGB, dst=pool[i].array)
        return pool[i].array, i

    def reader():
        try:
            image = first
            for frame_number in range(0, n_frames):
                if stop.is_set():
                    break
                if frame_number != 0:
                    success, image = vidcap.read()
                    if not success:
                        break
                if every is not None and every > 0 and (frame_number % every) != 0:
                    continue
                if wanted is not None:
                    if frame_number > last_wanted:
                        break
                    if frame_number not in wanted:
                        continue
                if image.shape != first.shape:  # a stream that changes resolution: leave the pinned pool
                    q.put((frame_number, cv2.cvtColor(image, cv2.COLOR_BGR2RGB), -1))
                else:
                    rgb, slot = to_rgb(image)
                    q.put((frame_number, rgb, slot))
        finally:
            q.put(None)

    th = threading.Thread(target=reader, daemon=True)
    th.start()
    frame_filenames, results = [], []
    try:
        done = False
        while not done:
            frames, ids, slots = [], [], []
            while len(frames) < batch_size:
                item = q.get()
                if item is None:
                    done = True
                    break
                frames.append(item[1])
                ids.append(frame_number_to_filename(item[0]))
                slots.append(item[2])
            if frames:
                res = detector.generate_detections_one_batch(frames, ids, detection_threshold=detection_threshold,
                                                             image_size=image_size)
                frame_filenames.extend(ids)
                results.extend(res)
                for sl in slots:
                    if sl >= 0:
                        free_q.put(sl)
    finally:
        stop.set()
        while th.is_alive():  # unblock a reader stuck on a full queue (or waiting for a free pinned buffer)
            try:
                item = q.get_nowait()
                if item is not None and free_q is not None and item[2] >= 0:
                    free_q.put(item[2])
            except queue.Empty:
                time.sleep(0.001)
        vidcap.release()
    if len(frame_filenames) == 0:
        if allow_empty_videos:
            print('Warning: found no frames in file {}'.format(input_video_file))
        else:
            raise Exception('Error: found no frames in file {}'.format(input_video_file))
    return {'frame_filenames': frame_filenames, 'frame_rate': frame_rate, 'results': results}


def video_results_to_md_format(video_filename, frame_rate, frame_results, confidence_threshold=None):
    """process_video.py:214-258: one "images" entry for a video; frame_results is the 'results' list above, or a dict with a
    'failure' key for a video that could not be read."""
    import re
    im = {'file': video_filename, 'frame_rate': frame_rate, 'frames_processed': []}
    if isinstance(frame_results, dict):
        im['failure'] = frame_results['failure']
        im['detections'] = None
        return im
    im['detections'] = []
    for r in frame_results:
        frame_number = int(re.search(r'frame(\d+)\.jpg', os.path.basename(r['file'])).group(1))
        assert frame_number not in im['frames_processed'], 'Received the same frame twice for video {}'.format(video_filename)
        im['frames_processed'].append(frame_number)
        for det in (r.get('detections') or []):
            if confidence_threshold is not None and det['conf'] < confidence_threshold:
                continue
            d = dict(det)
            d['frame_number'] = frame_number
            im['detections'].append(d)
    im['frames_processed'] = sorted(im['frames_processed'])
    return im


def find_videos(folder, recursive=True):
    out = []
    for root, dirs, files in os.walk(folder):
        out.extend(os.path.join(root, f) for f in files if f.lower().endswith(VIDEO_EXTENSIONS))
        if not recursive:
            break
    return sorted(out)


def process_videos(detector, input_video_file, output_json_file=None, frame_sample=None, time_sample=None, batch_size=32,
                   json_confidence_threshold=0.005, recursive=True, detector_file=None):
    """process_video.process_videos (process_video.py:123-280) with batched frames; returns the MD-format dict."""
    if frame_sample is not None and time_sample is not None:
        raise ValueError('frame_sample and time_sample are mutually exclusive')
    every = (-1 * time_sample) if time_sample is not None else frame_sample
    if os.path.isfile(input_video_file):
        folder = os.path.dirname(input_video_file)
        videos = [input_video_file]
    else:
        assert os.path.isdir(input_video_file), '{} is neither a file nor a folder'.format(input_video_file)
        folder = input_video_file
        videos = find_videos(folder, recursive)
    images = []
    for v in videos:
        rel = os.path.relpath(v, folder).replace('\\', '/') if folder else os.path.basename(v)
        try:
            r = run_detector_on_video(detector, v, every_n_frames=every, batch_size=batch_size,
                                      detection_threshold=json_confidence_threshold, allow_empty_videos=True)
            for fr in r['results']:
                fr['file'] = rel + '/' + fr['file']
            images.append(video_results_to_md_format(rel, r['frame_rate'], r['results']))
        except Exception as e:
            print('Video {} failed: {}'.format(v, e))
            images.append(video_results_to_md_format(rel, -1, {'failure': FAILURE_VIDEO_OPEN}))
    # process_video.py:261-265: the records go through run_detector_batch.write_results_to_file
    from . import run_batch
    if output_json_file is not None:
        run_batch.write_results(images, output_json_file, relative_path_base=None, detector_file=detector_file)
    out = run_batch.results_document(images, None, detector_file)
    return out
