#!/usr/bin/env python
"""
BASELINE.json config 1, for real: the reference's OWN batch driver
    megadetector.detection.run_detector_batch.load_and_run_detector_batch   (run_detector_batch.py:1062)
    -> run_detector.load_detector                                           (run_detector.py:601-681, with the 6-line
                                                                             INTEGRATION.md hook applied as a monkeypatch)
    -> write_results_to_file                                                (run_detector_batch.py:1546)
    -> postprocessing.validate_batch_results
on the repo's three sample JPEGs (copied to tests/golden/config1_images/).  The unmodified reference package is imported
from baseline/_ref (tools/install_reference.sh) with two import-time stubs (tests/stubs/).

    python tools/run_config1.py --engine b200   --out b200.json   [--batch-size 4] [--image-queue] [--preprocess-on-queue]
    python tools/run_config1.py --engine oracle --out oracle.json                 # the CPU arm (oracle port of the path)

The "model file" is a synthetic yolov5-format .pt (no MDv5a offline): megadetector_b200/synth_checkpoint.py.
"""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'baseline', '_ref'), os.path.join(ROOT, 'tests', 'stubs')):
    if p not in sys.path:
        sys.path.insert(0, p)

DEFAULT_SPEC = 'synthetic://yolov5x6?seed=0&profile=contractive'
IMAGES = ['nacti.jpg', 'idaho-camera-traps.jpg', 'orinoquia-thumb-web.jpg']


def install_hook():
    """The INTEGRATION.md hook, applied from outside: `detector_options['engine']` selects the detector class for `.pt`
    files; everything else falls through to the reference's own load_detector."""
    from megadetector.detection import run_detector, run_detector_batch
    original = run_detector.load_detector

    def load_detector(model_file, force_cpu=False, force_model_download=False, detector_options=None, verbose=False):
        engine = str((detector_options or {}).get('engine', os.environ.get('MD_ENGINE', ''))).lower()
        if model_file.endswith('.pt') and engine == 'b200':
            from megadetector_b200.detector import B200Detector
            detector_options['use_model_native_classes'] = run_detector.USE_MODEL_NATIVE_CLASSES
            return B200Detector(model_file, detector_options, verbose=verbose)
        if model_file.endswith('.pt') and engine == 'oracle':  # test infrastructure: the CPU arm
            from oracle.pt_detector import OraclePTDetector
            return OraclePTDetector(model_file, detector_options, verbose=verbose)
        return original(model_file, force_cpu=force_cpu, force_model_download=force_model_download,
                        detector_options=detector_options, verbose=verbose)

    run_detector.load_detector = load_detector
    run_detector_batch.load_detector = load_detector
    try:  # process_video binds the name at import time too (process_video.py:158 `detector = load_detector(...)`)
        from megadetector.detection import process_video
        process_video.load_detector = load_detector
    except Exception:
        pass
    return run_detector, run_detector_batch


def run(engine, out, spec=DEFAULT_SPEC, batch_size=4, image_queue=False, preprocess_on_queue=False, model_file=None,
        image_dir=None):
    run_detector, rdb = install_hook()
    from megadetector.postprocessing import validate_batch_results as vbr
    image_dir = image_dir or os.path.join(ROOT, 'tests', 'golden', 'config1_images')
    files = [os.path.join(image_dir, f) for f in IMAGES]
    tmp = None
    if model_file is None:
        from megadetector_b200.synth_checkpoint import write_synthetic_checkpoint
        tmp = tempfile.mkdtemp(prefix='mdb200_config1_')
        model_file = os.path.join(tmp, 'md_synthetic.pt')
        write_synthetic_checkpoint(spec, model_file)
    opts = {'engine': engine, 'max_batch': max(1, batch_size)}
    results = rdb.load_and_run_detector_batch(model_file, files, confidence_threshold=0.005, quiet=True,
                                              detector_options=opts, batch_size=batch_size,
                                              use_image_queue=image_queue, preprocess_on_image_queue=preprocess_on_queue,
                                              use_threads_for_queue=True, loader_workers=2)
    rdb.write_results_to_file(results, out, relative_path_base=image_dir, detector_file='md_synthetic.pt')
    options = vbr.ValidateBatchResultsOptions()
    options.raise_errors = True
    vbr.validate_batch_results(out, options)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--engine', required=True, choices=['b200', 'oracle'])
    ap.add_argument('--out', required=True)
    ap.add_argument('--spec', default=DEFAULT_SPEC)
    ap.add_argument('--batch-size', type=int, default=4)
    ap.add_argument('--image-queue', action='store_true')
    ap.add_argument('--preprocess-on-queue', action='store_true')
    ap.add_argument('--model-file', default=None)
    args = ap.parse_args()
    print(run(args.engine, args.out, args.spec, args.batch_size, args.image_queue, args.preprocess_on_queue, args.model_file))


if __name__ == '__main__':
    main()
