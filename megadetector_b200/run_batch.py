"""
In-process multi-GPU batch driver: `python -m megadetector_b200.run_batch MODEL IMAGE_DIR OUT.json --gpus N`.

The reference scales out by hand: the notebook splits the file list into N contiguous chunks, writes N shell scripts with
CUDA_VISIBLE_DEVICES=i, runs N copies of the CLI and merges the N JSON files
(/root/reference/notebooks/manage_local_batch.py:496-547,933-975; the author's TODO for an in-process `--gpus N`:
/root/reference/TODO.md:203-205).  This module does the same inside one command: the file list is split with the
reference's own chunking rule (sharding.split_list_into_n_chunks == ct_utils.split_list_into_n_chunks 'greedy'), one worker
process per GPU owns one `B200Detector` (one `mdb_engine`, images independent end to end, NO collective on the data path),
decodes its images with a thread pool, runs them in batches and returns its records; the parent merges them with the
notebook's uniqueness check and writes one results file in the reference's format (run_detector_batch.py:1546-1700,
format_version 1.6: info incl. detector_metadata / detection_categories / images).

Checkpoint / resume use the reference's own file format and flags (run_detector_batch.py:1465-1520, 1983-2003, 2095-2118,
2179-2181): a JSON dict {"checkpoint": [records with absolute file names]} rewritten every --checkpoint_frequency images
(previous file kept as <path>_tmp while writing), `--resume_from_checkpoint auto|PATH` skips the images already in it, the
checkpoint is deleted after the results file has been written.  A checkpoint written by the reference CLI resumes here and
vice versa.  With N workers the parent owns the checkpoint: workers stream their records back batch by batch.

The result file is diff-equal to what N runs of `run_detector_batch` behind the INTEGRATION.md hook produce, because every
record comes from the same `B200Detector.generate_detections_one_batch` call and images never interact.
"""
import argparse
import json
import os
import shutil
import sys
import time
from datetime import datetime
from concurrent.futures import ThreadPoolExecutor

from . import sharding

IMAGE_EXTENSIONS = ('.jpg', '.jpeg', '.gif', '.png', '.tif', '.tiff', '.bmp')  # path_utils.IMG_EXTENSIONS
DEFAULT_OUTPUT_CONFIDENCE_THRESHOLD = 0.005  # run_detector.py DEFAULT_OUTPUT_CONFIDENCE_THRESHOLD
CURRENT_FORMAT_VERSION = '1.6'               # run_detector_batch.py:119

# The MDv5 part of run_detector.known_models / model_string_to_model_version (run_detector.py:88-216): what
# write_results_to_file records as info.detector_metadata for the models this engine runs.  Order matters (first match wins).
_MDV5_FILENAME_STRINGS = (('mdv5a', 'v5a.0.1'), ('mdv5b', 'v5b.0.1'), ('v5a.0.0', 'v5a.0.1'), ('v5b.0.0', 'v5b.0.1'),
                          ('v5a.0.1', 'v5a.0.1'), ('v5b.0.1', 'v5b.0.1'), ('mdv5', 'v5a.0.1'), ('md5', 'v5a.0.1'),
                          ('default', 'v5a.0.1'), ('megadetector', 'v5a.0.1'))
_MDV5_MD5 = {'v5a.0.0': 'ec1d7603ec8cf642d6e0cd008ba2be8c', 'v5b.0.0': 'bc235e73f53c5c95e66ea0d1b2cbf542',
             'v5a.0.1': '60f8e7ec1308554df258ed1f4040bc4f', 'v5b.0.1': 'f17ed6fedfac2e403606a08c89984905'}


def detector_metadata_from_filename(detector_file):
    """get_detector_version_from_filename + get_detector_metadata_from_version_string (run_detector.py:276-350) for the
    MDv5 family; any other name gets the reference's 'unknown' block."""
    fn = os.path.basename(detector_file).lower()
    version = next((v for key, v in _MDV5_FILENAME_STRINGS if key in fn), 'unknown')
    if version not in _MDV5_MD5:
        return {'megadetector_version': 'unknown', 'typical_detection_threshold': 0.2, 'conservative_detection_threshold': 0.1}
    return {'url': 'https://github.com/agentmorris/MegaDetector/releases/download/v5.0/md_{}.pt'.format(version),
            'typical_detection_threshold': 0.2, 'conservative_detection_threshold': 0.05, 'image_size': 1280,
            'model_type': 'yolov5', 'normalized_typical_inference_speed': 1.0, 'md5': _MDV5_MD5[version],
            'megadetector_version': version}


def find_images(folder, recursive=True):
    """path_utils.find_images semantics: case-insensitive extension match, sorted."""
    out = []
    if recursive:
        for root, _, files in os.walk(folder):
            out.extend(os.path.join(root, f) for f in files if f.lower().endswith(IMAGE_EXTENSIONS))
    else:
        out = [os.path.join(folder, f) for f in os.listdir(folder) if f.lower().endswith(IMAGE_EXTENSIONS)]
    return sorted(out)


def resolve_image_list(image_file, recursive=True):
    """The reference CLI's `image_file` argument (run_detector_batch.py:2005-2040): a folder, a .json list of paths, a .txt
    list (one path per line) or a single image."""
    if os.path.isdir(image_file):
        return find_images(image_file, recursive)
    if os.path.isfile(image_file) and image_file.endswith('.json'):
        with open(image_file) as f:
            return list(json.load(f))
    if os.path.isfile(image_file) and image_file.endswith('.txt'):
        with open(image_file) as f:
            return [ln.strip() for ln in f if ln.strip()]
    if os.path.isfile(image_file) and image_file.lower().endswith(IMAGE_EXTENSIONS):
        return [image_file]
    raise ValueError('image_file specified is not a directory, a json list, a txt list, or an image file')


def load_image(path):
    """visualization_utils.load_image (visualization_utils.py:103-176,306-323): PIL, RGB, EXIF orientation applied."""
    from PIL import Image
    image = Image.open(path)
    if image.mode not in ('RGBA', 'RGB', 'L', 'I;16'):
        raise AttributeError('Image {} uses unsupported mode {}'.format(path, image.mode))
    if image.mode in ('RGBA', 'L'):
        image = image.convert(mode='RGB')
    try:
        exif = image._getexif()
        orientation = exif.get(274, None)
        rotations = {3: 180, 6: 270, 8: 90}
        if orientation is not None and orientation != 1 and orientation in rotations:
            image = image.rotate(rotations[orientation], expand=True)
    except Exception:
        pass
    image.load()
    return image


class _Sink(list):
    """The list a worker collects its records in; `flush()` hands the records added since the last call to `on_batch`."""

    def __init__(self, on_batch=None):
        super().__init__()
        self._on_batch, self._sent = on_batch, 0

    def flush(self):
        if self._on_batch is not None and len(self) > self._sent:
            self._on_batch(self[self._sent:])
        self._sent = len(self)


def process_files(detector, files, batch_size, confidence_threshold, loader_threads=8, quiet=True, gpu_decode=False,
                  on_batch=None):
    """One worker's loop: decode with a thread pool one batch ahead, detect, apply the output threshold
    (run_detector_batch.py:724-767).  gpu_decode: `.jpg` / `.jpeg` files are read as bytes and decoded by nvJPEG on the GPU
    (B200Detector.generate_detections_from_jpeg_files); other formats keep the PIL path.  on_batch(records): called after
    every batch with that batch's records (the checkpoint stream)."""
    results = _Sink(on_batch)
    if gpu_decode and hasattr(detector, 'generate_detections_from_jpeg_files'):
        jpegs = [f for f in files if f.lower().endswith(('.jpg', '.jpeg'))]
        files = [f for f in files if not f.lower().endswith(('.jpg', '.jpeg'))]

        def read(f):
            try:
                with open(f, 'rb') as fh:
                    return (f, fh.read())
            except Exception:
                return (f, b'')

        jb = [jpegs[i:i + batch_size] for i in range(0, len(jpegs), batch_size)]
        with ThreadPoolExecutor(max_workers=loader_threads) as pool:
            pending = list(pool.map(read, jb[0])) if jb else None
            for bi in range(len(jb)):
                blobs = pending
                nxt = pool.map(read, jb[bi + 1]) if bi + 1 < len(jb) else None
                for r in detector.generate_detections_from_jpeg_files(blobs):
                    if r.get('detections') is None:
                        results.append({'file': r['file'], 'failure': r['failure']})
                    else:
                        r['detections'] = [d for d in r['detections'] if d['conf'] >= confidence_threshold]
                        results.append(r)
                results.flush()
                pending = list(nxt) if nxt is not None else None

    def load(f):
        try:
            return load_image(f)
        except Exception as e:
            if not quiet:
                print('Image {} cannot be loaded: {}'.format(f, e))
            return None

    batches = [files[i:i + batch_size] for i in range(0, len(files), batch_size)]
    with ThreadPoolExecutor(max_workers=loader_threads) as pool:
        pending = pool.map(load, batches[0]) if batches else None
        for bi, batch in enumerate(batches):
            images = list(pending)
            pending = pool.map(load, batches[bi + 1]) if bi + 1 < len(batches) else None
            ok = [(f, im) for f, im in zip(batch, images) if im is not None]
            for f, im in zip(batch, images):
                if im is None:
                    results.append({'file': f, 'failure': 'image access failure'})
            if ok:
                recs = detector.generate_detections_one_batch([im for _, im in ok], [f for f, _ in ok])
                for r in recs:
                    if 'failure' in r and r.get('detections') is None:
                        results.append({'file': r['file'], 'failure': r['failure']})
                        continue
                    r['detections'] = [d for d in r['detections'] if d['conf'] >= confidence_threshold]
                    results.append(r)
            results.flush()
    return list(results)


def write_checkpoint(checkpoint_path, results):
    """run_detector_batch.write_checkpoint (:1465-1494): {"checkpoint": results}; the previous file survives as *_tmp
    until the new one is complete."""
    tmp = None
    if os.path.isfile(checkpoint_path):
        tmp = checkpoint_path + '_tmp'
        shutil.copyfile(checkpoint_path, tmp)
    with open(checkpoint_path, 'w', newline='\n', encoding='utf-8') as f:
        json.dump({'checkpoint': results}, f, indent=1, default=str, ensure_ascii=True)
    if tmp is not None:
        try:
            os.remove(tmp)
        except Exception as e:
            print('Warning: error removing backup checkpoint file {}:\n{}'.format(tmp, e))


def load_checkpoint(checkpoint_path):
    """run_detector_batch.load_checkpoint (:1497-1520)."""
    print('Loading previous results from checkpoint file {}'.format(checkpoint_path))
    with open(checkpoint_path, 'r') as f:
        data = json.load(f)
    if 'checkpoint' not in data:
        raise ValueError('Checkpoint file {} is missing "checkpoint" field'.format(checkpoint_path))
    print('Restored {} entries from the checkpoint {}'.format(len(data['checkpoint']), checkpoint_path))
    return data['checkpoint']


def find_checkpoint(output_dir):
    """`--resume_from_checkpoint auto` (:1984-1997): the latest md_checkpoint*.json next to the output file."""
    found = sorted(fn for fn in os.listdir(output_dir) if fn.startswith('md_checkpoint') and fn.endswith('.json'))
    if not found:
        raise ValueError('resume_from_checkpoint set to "auto", but no checkpoints found in {}'.format(output_dir))
    if len(found) > 1:
        print('Warning: found {} checkpoints in {}, using the latest'.format(len(found), output_dir))
    return os.path.join(output_dir, found[-1])


def _worker(rank, world, model_file, files, batch_size, confidence_threshold, detector_options, factory, out_queue,
            gpu_decode=False, stream=False):
    try:
        opts = dict(detector_options or {})
        opts.setdefault('max_batch', batch_size)
        opts['device'] = 'cuda:{}'.format(rank)
        if factory is not None:
            mod, name = factory.rsplit(':', 1)
            import importlib
            detector = getattr(importlib.import_module(mod), name)(model_file, opts)
        else:
            from .detector import B200Detector
            detector = B200Detector(model_file, opts)
        mine = sharding.shard_for_rank(files, rank, world)
        t0 = time.time()
        on_batch = (lambda recs: out_queue.put((rank, list(recs), None, 'batch'))) if stream else None
        res = process_files(detector, mine, batch_size, confidence_threshold, gpu_decode=gpu_decode, on_batch=on_batch)
        out_queue.put((rank, None if stream else res, time.time() - t0, None))
        if hasattr(detector, 'close'):
            detector.close()
    except Exception as e:  # never leave the parent waiting
        import traceback
        out_queue.put((rank, None, 0.0, '{}\n{}'.format(e, traceback.format_exc())))


def run_sharded(model_file, files, n_gpus, batch_size=64, confidence_threshold=DEFAULT_OUTPUT_CONFIDENCE_THRESHOLD,
                detector_options=None, detector_factory=None, gpu_decode=False, checkpoint_path=None,
                checkpoint_frequency=-1, results=None):
    """
    Runs `files` on `n_gpus` GPUs of this box, one process per GPU.  Returns the merged records sorted by file.
    detector_factory: optional 'module:callable' (model_file, options) -> detector, for tests of the plumbing.
    results: records restored from a checkpoint; their files are skipped (load_and_run_detector_batch :1204, 1274-1279) and
    they are part of what is returned.  checkpoint_path + checkpoint_frequency > 0: the parent rewrites the checkpoint every
    time the number of finished images crosses a multiple of the frequency (:272-288).
    """
    import multiprocessing as mp
    results = list(results or [])
    already = set(r['file'] for r in results)
    todo = [f for f in files if f not in already]
    if len(todo) != len(files):
        print('Bypassing {} images that have already been processed'.format(len(files) - len(todo)))
    files = todo
    if not files:
        return sorted(results, key=lambda r: r['file']), []
    checkpointing = checkpoint_path is not None and checkpoint_frequency is not None and checkpoint_frequency > 0
    n_gpus = max(1, min(int(n_gpus), max(1, len(files))))
    ctx = mp.get_context('spawn')  # CUDA contexts do not survive fork
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, n_gpus, model_file, files, batch_size, confidence_threshold,
                                               detector_options, detector_factory, q, gpu_decode, checkpointing))
             for r in range(n_gpus)]
    for p in procs:
        p.start()
    parts, times = [[] for _ in range(n_gpus)], [0.0] * n_gpus
    done, n_new, last_written = 0, 0, 0
    while done < n_gpus:
        rank, res, dt, err = q.get()
        if err == 'batch':  # a worker's records of one batch (checkpoint stream)
            parts[rank].extend(res)
            n_new += len(res)
            if n_new // checkpoint_frequency > last_written // checkpoint_frequency:
                print('Writing a new checkpoint after having processed {} images since last restart'.format(n_new))
                write_checkpoint(checkpoint_path, results + [r for part in parts for r in part])
                last_written = n_new
            continue
        if err is not None:
            for p in procs:
                p.terminate()
            if checkpointing and n_new > last_written:  # keep what finished before the failure
                write_checkpoint(checkpoint_path, results + [r for part in parts for r in part])
            raise RuntimeError('worker {} failed: {}'.format(rank, err))
        if res is not None:
            parts[rank] = res
        times[rank] = dt
        done += 1
    for p in procs:
        p.join()
    merged = results + [r for part in parts for r in part]
    names = [r['file'] for r in merged]
    if len(set(names)) != len(names):  # manage_local_batch.py:947-958
        raise ValueError('duplicate images in the merged results')
    return sorted(merged, key=lambda r: r['file']), times


def results_document(results, relative_path_base=None, detector_file=None, include_max_conf=False):
    """The dict run_detector_batch.write_results_to_file (:1546-1662, default arguments) builds: images sorted by file,
    detections by descending confidence, failure records with detections None, info with format 1.6 + detector metadata."""
    images = []
    for r in results:
        r = dict(r)
        if relative_path_base is not None:
            r['file'] = os.path.relpath(r['file'], relative_path_base)
        r['file'] = r['file'].replace('\\', '/')
        if not include_max_conf:
            r.pop('max_detection_conf', None)
        if r.get('detections') is not None:  # :1634-1637: descending confidence (stable)
            r['detections'] = sorted(r['detections'], key=lambda d: d['conf'], reverse=True)
        if 'failure' in r:                   # :1639-1644
            assert r.get('detections') is None, 'Illegal failure/detection combination'
            r['detections'] = None
        images.append(r)
    images.sort(key=lambda r: r['file'])
    info = {'detection_completion_time': datetime.now().strftime('%Y-%m-%d %H:%M:%S'), 'format_version': CURRENT_FORMAT_VERSION}
    if detector_file is not None:
        info['detector'] = os.path.basename(detector_file)
        info['detector_metadata'] = detector_metadata_from_filename(detector_file)
    else:
        info['detector'] = 'unknown'
        info['detector_metadata'] = detector_metadata_from_filename('unknown')
    return {'images': images, 'detection_categories': {'1': 'animal', '2': 'person', '3': 'vehicle'}, 'info': info}


def write_results(results, output_file, relative_path_base=None, detector_file=None, include_max_conf=False):
    """run_detector_batch.write_results_to_file: same JSON file (ct_utils.write_json with force_str, :210-250)."""
    out = results_document(results, relative_path_base, detector_file, include_max_conf)
    try:
        os.makedirs(os.path.dirname(output_file), exist_ok=True)
    except Exception:
        pass
    with open(output_file, 'w', newline='\n', encoding='utf-8') as f:  # ct_utils.write_json(force_str=True) (:210-250)
        json.dump(out, f, indent=1, default=str, ensure_ascii=True)
    return output_file


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n\n')[0])
    ap.add_argument('detector_file')
    ap.add_argument('image_dir', metavar='image_file', help='folder, .json / .txt list of image paths, or one image')
    ap.add_argument('output_file')
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--batch_size', type=int, default=64)
    ap.add_argument('--threshold', type=float, default=DEFAULT_OUTPUT_CONFIDENCE_THRESHOLD)
    ap.add_argument('--recursive', action='store_true')
    ap.add_argument('--output_relative_filenames', action='store_true')
    ap.add_argument('--gpu_decode', action='store_true', help='decode .jpg files with nvJPEG on the GPU instead of PIL')
    # the reference CLI's checkpoint flags (run_detector_batch.py:1822-1848)
    ap.add_argument('--checkpoint_frequency', type=int, default=-1, help='write a checkpoint every N images; -1 disables')
    ap.add_argument('--checkpoint_path', type=str, default=None,
                    help='checkpoint file; default md_checkpoint_<timestamp>.json next to the output file')
    ap.add_argument('--resume_from_checkpoint', type=str, default=None,
                    help='checkpoint file to resume from, or "auto" for the latest md_checkpoint*.json next to the output file')
    ap.add_argument('--allow_checkpoint_overwrite', action='store_true')
    ap.add_argument('--include_max_conf', action='store_true', help='keep max_detection_conf in the output (off by default)')
    args = ap.parse_args(argv)
    output_dir = os.path.dirname(os.path.abspath(args.output_file))
    os.makedirs(output_dir, exist_ok=True)
    restored = []
    if args.resume_from_checkpoint is not None:
        ckpt = find_checkpoint(output_dir) if args.resume_from_checkpoint == 'auto' else args.resume_from_checkpoint
        restored = load_checkpoint(ckpt)
    checkpoint_path = None
    if args.checkpoint_frequency != -1:
        checkpoint_path = args.checkpoint_path or os.path.join(
            output_dir, 'md_checkpoint_{}.json'.format(datetime.now().strftime('%Y%m%d%H%M%S')))
        if not args.allow_checkpoint_overwrite and args.resume_from_checkpoint is None:
            assert not os.path.isfile(checkpoint_path), \
                'Checkpoint path {} already exists, delete or move it before re-using the same checkpoint path, ' \
                'or specify --allow_checkpoint_overwrite'.format(checkpoint_path)
        print('The checkpoint file will be written to {}'.format(checkpoint_path))
    elif args.checkpoint_path is not None:
        print('Warning: checkpointing disabled because checkpoint_frequency is -1, but a checkpoint path was specified')
    # file names in a checkpoint are absolute (:1980-1981); relative names are made when the results are written
    if args.output_relative_filenames:
        assert os.path.isdir(args.image_dir), 'Could not find folder {}, must supply a folder when ' \
            '--output_relative_filenames is set'.format(args.image_dir)
    files = resolve_image_list(args.image_dir, args.recursive)
    if restored or checkpoint_path:
        files = [os.path.abspath(f) for f in files]
    t0 = time.time()
    results, times = run_sharded(args.detector_file, files, args.gpus, args.batch_size, args.threshold,
                                 gpu_decode=args.gpu_decode, checkpoint_path=checkpoint_path,
                                 checkpoint_frequency=args.checkpoint_frequency, results=restored)
    write_results(results, args.output_file, args.image_dir if args.output_relative_filenames else None, args.detector_file,
                  include_max_conf=args.include_max_conf)
    if checkpoint_path and os.path.isfile(checkpoint_path):
        os.remove(checkpoint_path)
        print('Deleted checkpoint file {}'.format(checkpoint_path))
    print('Processed {} images on {} GPUs in {:.1f} s (per-GPU detector time: {})'.format(
        len(files), args.gpus, time.time() - t0, ', '.join('{:.1f}'.format(t) for t in times)))
    return 0


if __name__ == '__main__':
    sys.exit(main())
