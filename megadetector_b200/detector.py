"""
`B200Detector`: the drop-in for `megadetector.detection.pytorch_detector.PTDetector`
(/root/reference/megadetector/detection/pytorch_detector.py:739-1482) backed by libmdb200 (hand-written
sm_100a CUDA behind the C ABI of include/mdb200.h).

Same constructor signature, same three methods, same result records, same failure strings, same exceptions
for API misuse -- so `run_detector.load_detector()` can return it and
`run_detector_batch` / `process_video` drive it unchanged.

What runs where:
  * raw images (PIL / ndarray)   -> mdb_infer_batch:  K1 letterbox .. K8 rescale, all on the GPU
  * dicts from producer workers  -> mdb_infer_preprocessed (the letterboxed u8 image is uploaded as is)
  * `preprocess_image()` itself  -> host code (cv2), exactly what the reference's producer processes do
    (run_detector_batch.py:157-183): it exists for the image-queue workers, which must not touch the GPU
    (pytorch_detector.py:785-789,850-851).  It is not a fallback for K1.
  * record formatting            -> vectorised numpy replica of pytorch_detector.py:1384-1411 (fp32 -> double ->
    decimal truncation, op for op)

Compatibility modes: 'classic' (and any mode string containing it) and the non-classic ("modern") behaviour of
pytorch_detector.py:1036-1101,1318-1321,1372-1397: INTER_AREA / INTER_LINEAR pre-resize of the long side, padding to
ceil(side / stride + 0.5) * stride, NMS IoU 0.6, scale_coords with ratio_pad, round() instead of truncation.
`augment=True` runs yolov5's test-time augmentation (three scaled / flipped passes) inside the engine (mdb_set_augment).
Not supported (raises NotImplementedError, never a silent approximation): `force_cpu` (there is no CPU path).
"""

import gc
import math

import numpy as np

from . import _capi
from . import checkpoint

# run_detector.py:55-67 (kept identical; the reference module is not importable on a bare GPU box)
FAILURE_INFER = 'inference failure'
FAILURE_IMAGE_OPEN = 'image access failure'
CONF_DIGITS = 3
COORD_DIGITS = 4
DEFAULT_DETECTOR_LABEL_MAP = {'1': 'animal', '2': 'person', '3': 'vehicle'}

default_compatibility_mode = 'classic'


def _parse_bool(v):
    # ct_utils.parse_bool_string (utils/ct_utils.py:1000-1031) semantics for the values that occur in practice
    if isinstance(v, bool):
        return v
    s = str(v).strip().lower()
    if s in ('true', '1'):
        return True
    if s in ('false', '0'):
        return False
    raise ValueError('Cannot parse bool from string {}'.format(v))


def _truncate(x, precision):
    """ct_utils.truncate_float (utils/ct_utils.py:82-103) on float64 arrays: floor(x * 10^p) / 10^p."""
    f = float(10 ** precision)
    return np.floor(x * f) / f


_mdbfmt = None


def _formatter():
    """The C record formatter (csrc/pyformat.c), built in-tree on first use like the CUDA library."""
    global _mdbfmt
    if _mdbfmt is None:
        import importlib
        from .build import build_formatter
        build_formatter()
        _mdbfmt = importlib.import_module('megadetector_b200._mdbfmt')
    return _mdbfmt


def format_detections(det, count, src_h, src_w, use_model_native_classes=False, classic=True):
    """
    pytorch_detector.py:1367-1411: det rows = x1,y1,x2,y2 (rounded original pixels, fp32), conf, cls, sorted by conf
    descending.  Returns (list of dicts in ascending conf order, max_conf).  classic: truncate_float (:1392-1394), otherwise
    Python round() (:1395-1397).  Runs in C (csrc/pyformat.c: 0.15 us per detection instead of 0.65); the numpy statement
    of the same arithmetic below is what tests/test_detector_host.py holds it against.
    """
    n = int(count)
    if n == 0:
        return [], 0.0
    d = np.ascontiguousarray(det[:n], dtype=np.float32)
    return _formatter().format_detections(d, n, int(src_h), int(src_w), bool(use_model_native_classes), bool(classic))


def format_detections_numpy(det, count, src_h, src_w, use_model_native_classes=False, classic=True):
    """The same record formatting in numpy (the executable specification of csrc/pyformat.c)."""
    n = int(count)
    if n == 0:
        return [], 0.0
    d = det[:n][::-1]  # `for *xyxy, conf, cls in reversed(det)` (:1384)
    x1, y1, x2, y2 = d[:, 0], d[:, 1], d[:, 2], d[:, 3]  # float32
    two = np.float32(2.0)
    # yolov5 xyxy2xywh in float32, then `/ gn` with gn = [W,H,W,H] (int64 -> float32 division) (:1367,:1389)
    gw, gh = np.float32(src_w), np.float32(src_h)
    xc = ((x1 + x2) / two) / gw
    yc = ((y1 + y2) / two) / gh
    bw = (x2 - x1) / gw
    bh = (y2 - y1) / gh
    # .tolist() -> Python doubles; ct_utils.convert_yolo_to_xywh in double (utils/ct_utils.py:255-270)
    xc, yc, bw, bh = (a.astype(np.float64) for a in (xc, yc, bw, bh))
    x_min = xc - bw / 2.0
    y_min = yc - bh / 2.0
    if classic:
        box = np.stack([_truncate(x_min, COORD_DIGITS), _truncate(y_min, COORD_DIGITS), _truncate(bw, COORD_DIGITS),
                        _truncate(bh, COORD_DIGITS)], 1)
        conf = _truncate(d[:, 4].astype(np.float64), CONF_DIGITS)
    else:
        # ct_utils.round_float == Python's round() (correctly rounded decimal, half-to-even): not np.round
        box = np.array([[round(v, COORD_DIGITS) for v in row] for row in np.stack([x_min, y_min, bw, bh], 1).tolist()],
                       np.float64).reshape(-1, 4)
        conf = np.array([round(v, CONF_DIGITS) for v in d[:, 4].astype(np.float64).tolist()], np.float64)
    cls = d[:, 5].astype(np.int64)
    if not use_model_native_classes:
        cls = cls + 1
        bad = (cls < 1) | (cls > 3)
        if bad.any():
            raise KeyError('{} is not a valid class.'.format(int(cls[bad][0])))
    box_l = box.tolist()
    conf_l = conf.tolist()
    cls_l = cls.tolist()
    dets = [{'category': str(c), 'conf': cf, 'bbox': b} for c, cf, b in zip(cls_l, conf_l, box_l)]
    return dets, max(0.0, max(conf_l))


class B200Detector:
    """Runs a YOLOv5-P6 MegaDetector model on one B200 through libmdb200."""

    def __init__(self, model_path, detector_options=None, verbose=False):
        if verbose:
            print('Initializing B200Detector (verbose)')
        force_cpu = False
        use_model_native_classes = False
        compatibility_mode = default_compatibility_mode
        max_batch = 32
        device = None
        if detector_options is not None:
            if 'force_cpu' in detector_options:
                force_cpu = _parse_bool(detector_options['force_cpu'])
            if 'use_model_native_classes' in detector_options:
                use_model_native_classes = _parse_bool(detector_options['use_model_native_classes'])
            if detector_options.get('compatibility_mode') is not None:
                compatibility_mode = detector_options['compatibility_mode']
            if 'max_batch' in detector_options:
                max_batch = int(detector_options['max_batch'])
            if detector_options.get('device') is not None:
                device = detector_options['device']
        preprocess_only = bool(detector_options and detector_options.get('preprocess_only'))
        if compatibility_mode == 'default':
            compatibility_mode = default_compatibility_mode
        if verbose or (not preprocess_only):
            print('Loading PT detector with compatibility mode {}'.format(compatibility_mode))

        self.model_path = model_path
        self.model_metadata = checkpoint.read_metadata_from_megadetector_model_file(model_path)
        if self.model_metadata is not None and 'image_size' in self.model_metadata:
            self.default_image_size = self.model_metadata['image_size']
            print('Loaded image size {} from model metadata'.format(self.default_image_size))
        else:
            print('No image size available in model metadata, defaulting to 1280')
            self.default_image_size = 1280
        self.device = 'cpu'
        self.printed_image_size_warning = False
        self.use_model_native_classes = use_model_native_classes
        self.compatibility_mode = compatibility_mode
        self.classic = 'classic' in compatibility_mode
        self.use_ceil_for_resize = 'use_ceil_for_resize' in compatibility_mode
        self.letterbox_stride = 64 if self.default_image_size == 1280 else 32
        print('Using model stride: {}'.format(self.letterbox_stride))
        self.half_precision = False
        self.engine = None
        self._cfg = None
        self._current_image_size = self.default_image_size
        if preprocess_only:
            # geometry helper only; no weights, no GPU (pytorch_detector.py:785-789,850-851)
            self._cfg = _capi.make_config(img_size=self.default_image_size, stride=self.letterbox_stride,
                                          compat_mode=0 if self.classic else 1, use_ceil_for_resize=self.use_ceil_for_resize)
            return

        if force_cpu:
            raise NotImplementedError('B200Detector has no CPU path (force_cpu was requested)')
        ordinal = 0
        if device is not None:
            print('Using caller-specified device: {}'.format(device))
            s = str(device)
            if not s.startswith('cuda'):
                raise NotImplementedError('B200Detector runs on CUDA devices only (got {})'.format(device))
            ordinal = int(s.split(':')[1]) if ':' in s else 0
        self.device = 'cuda:{}'.format(ordinal)
        # AddaxAI depends on this printout, don't remove it (pytorch_detector.py:884-885)
        print('PTDetector using device {}'.format(str(self.device).lower()))

        weights, arch = checkpoint.load_model(model_path)
        if self.letterbox_stride != 64:
            raise NotImplementedError('only stride-64 (YOLOv5-P6, 1280 px) models are supported')
        self._cfg = _capi.make_config(device=ordinal, max_batch=max_batch, img_size=self.default_image_size,
                                      stride=self.letterbox_stride, letterbox_auto=True, letterbox_scaleup=True,
                                      iou_thres=0.45 if self.classic else 0.6, max_det=300, ch=arch['ch'],
                                      n_backbone=arch['n_backbone'], n_head=arch['n_head'], nc=arch['nc'],
                                      anchors_px=arch.get('anchors_px'), compat_mode=0 if self.classic else 1,
                                      use_ceil_for_resize=self.use_ceil_for_resize)
        self.engine = _capi.Engine(self._cfg)
        self.engine.load_weights(weights)
        self.nc = arch['nc']

    # ------------------------------------------------------------------------------------------
    def preprocess_image(self, img_original, image_id='unknown', image_size=None, verbose=False):
        """pytorch_detector.py:964-1119: host-side preprocessing for producer workers (both compatibility modes)."""
        import cv2
        result = {'file': image_id}
        img_original_pil = None
        if not isinstance(img_original, np.ndarray):
            img_original_pil = img_original
            img_original = np.asarray(img_original)
        scaling_shape = img_original.shape
        if image_size is not None:
            assert isinstance(image_size, int)
            if not self.printed_image_size_warning:
                print('Using user-supplied image size {}'.format(image_size))
                self.printed_image_size_warning = True
        else:
            image_size = self.default_image_size
            self.printed_image_size_warning = False
        cfg = _capi.make_config(img_size=image_size, stride=self.letterbox_stride, compat_mode=0 if self.classic else 1,
                                use_ceil_for_resize=self.use_ceil_for_resize)
        g = _capi.letterbox_geometry(cfg, img_original.shape[0], img_original.shape[1])
        img = img_original
        if (img.shape[1], img.shape[0]) != tuple(g['new_unpad']):
            shrink = (not self.classic) and g['new_unpad'][0] < img.shape[1]
            img = cv2.resize(img, tuple(g['new_unpad']), interpolation=cv2.INTER_AREA if shrink else cv2.INTER_LINEAR)
        if not self.classic:
            img_original = img  # the reference keeps the pre-resized image under this key (:1060-1062,1112)
        img = cv2.copyMakeBorder(img, g['top'], g['bottom'], g['left'], g['right'], cv2.BORDER_CONSTANT,
                                 value=(114, 114, 114))
        result['img_processed'] = img
        result['img_original'] = img_original
        result['img_original_pil'] = img_original_pil
        result['target_shape'] = image_size if self.classic else np.array(g['out_hw'])
        result['scaling_shape'] = scaling_shape
        result['letterbox_ratio'] = g['ratio']
        result['letterbox_pad'] = g['pad']
        return result

    # ------------------------------------------------------------------------------------------
    def generate_detections_one_batch(self, img_original, image_id=None, detection_threshold=0.00001, image_size=None,
                                      augment=False, verbose=False):
        """pytorch_detector.py:1124-1252; returns a list of result dicts aligned with the inputs."""
        # API misuse raises the reference's ValueErrors with the reference's messages (pytorch_detector.py:1156-1182)
        if not isinstance(img_original, list):
            raise ValueError('img_original must be a list for batch processing')
        if len(img_original) == 0:
            return []
        dict_mode = isinstance(img_original[0], dict)
        if not dict_mode:
            if not isinstance(image_id, list):
                raise ValueError('image_id must be a list when img_original contains PIL/numpy images' if image_id is None
                                 else 'image_id must be a list for batch processing')
            if len(image_id) != len(img_original):
                raise ValueError('Length mismatch: img_original has {} items, image_id has {} items'.format(
                    len(img_original), len(image_id)))
        for i, img in enumerate(img_original):
            if isinstance(img, dict) != dict_mode:
                raise ValueError('Mixed input types in batch: item {} is {} a dict, but item 0 is {} a dict'.format(
                    i, *(('not', '') if dict_mode else ('', 'not'))).replace('is  a', 'is a'))
        if self.engine is None:
            raise RuntimeError('this B200Detector was created with preprocess_only; it cannot run inference')
        if detection_threshold is None:
            detection_threshold = 0.0
        size = self.default_image_size if image_size is None else image_size
        if image_size is not None:
            assert isinstance(image_size, int)
            if not self.printed_image_size_warning:
                print('Using user-supplied image size {}'.format(image_size))
                self.printed_image_size_warning = True
        else:
            self.printed_image_size_warning = False
        if size != self._current_image_size:
            self.engine.set_image_size(size)
            self._current_image_size = size

        n = len(img_original)
        results = [None] * n
        ids = [None] * n
        arrays = [None] * n
        src_shapes = [None] * n
        gain_pads = [None] * n
        for i, img in enumerate(img_original):
            try:
                if dict_mode:
                    ids[i] = img['file']
                    arrays[i] = np.ascontiguousarray(img['img_processed'], dtype=np.uint8)
                    src_shapes[i] = tuple(img['scaling_shape'][:2])
                    if not self.classic:  # ratio_pad of scale_coords (:1372-1381)
                        gain_pads[i] = (img['img_original'].shape[0] / img['scaling_shape'][0],
                                        img['letterbox_pad'][0], img['letterbox_pad'][1])
                else:
                    ids[i] = image_id[i]
                    a = img if isinstance(img, np.ndarray) else np.asarray(img)
                    if a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.uint8:
                        raise ValueError('expected an HxWx3 uint8 image, got shape {} dtype {}'.format(a.shape, a.dtype))
                    if a.strides[2] != 1 or a.strides[1] != 3:
                        a = np.ascontiguousarray(a)
                    arrays[i] = a
                    src_shapes[i] = (a.shape[0], a.shape[1])
            except Exception as e:
                print('Warning: preprocessing failed for image {}: {}'.format(
                    image_id[i] if image_id else f'index_{i}', str(e)))
                ids[i] = ids[i] if ids[i] is not None else (image_id[i] if image_id else f'index_{i}')
                results[i] = {'file': ids[i], 'detections': None, 'failure': FAILURE_IMAGE_OPEN}
        good = [i for i in range(n) if results[i] is None]

        def finish(j, i, det, cnt, status):
            if status[j] == _capi.MDB_IMG_BAD_INPUT:
                results[i] = {'file': ids[i], 'detections': None, 'failure': FAILURE_IMAGE_OPEN}
            elif status[j] != _capi.MDB_IMG_OK:
                if verbose:
                    print('Warning: batch inference failed for image {}: {}'.format(ids[i], self.engine.last_error()))
                results[i] = {'file': ids[i], 'detections': None, 'failure': FAILURE_INFER}
            else:
                try:
                    dets, max_conf = format_detections(det[j], cnt[j], src_shapes[i][0], src_shapes[i][1],
                                                       self.use_model_native_classes, self.classic)
                    results[i] = {'file': ids[i], 'detections': dets, 'max_detection_conf': max_conf}
                except KeyError as e:
                    # the reference raises inside _process_batch_group and records an inference failure (:1236-1249)
                    print('Warning: batch inference failed for image {}: {}'.format(ids[i], str(e)))
                    results[i] = {'file': ids[i], 'detections': None, 'failure': FAILURE_INFER}

        # A batch creates ~20 k small dicts/lists that all survive; with the cyclic collector enabled CPython
        # re-traverses them over and over (measured: 35 ms vs 8 ms per 64 images).  They are acyclic, so the
        # collector is paused while the records are built and restored afterwards.
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            self.engine.set_augment(bool(augment))  # yolov5 _forward_augment: 3 scaled / flipped passes (:1313)
            self._infer_and_format(good, dict_mode, arrays, src_shapes, detection_threshold, finish, results, ids, gain_pads)
        finally:
            self.engine.set_augment(False)
            if gc_was_enabled:
                gc.enable()
        return results

    def _infer_and_format(self, good, dict_mode, arrays, src_shapes, detection_threshold, finish, results, ids, gain_pads=None):
        if good:
            try:
                if dict_mode:
                    gp = None if self.classic else [gain_pads[i] for i in good]
                    det, cnt, _, status = self.engine.infer_preprocessed([arrays[i] for i in good],
                                                                         [src_shapes[i] for i in good],
                                                                         detection_threshold, want_anchors=False, gain_pad=gp)
                    for j, i in enumerate(good):
                        finish(j, i, det, cnt, status)
                else:
                    # records of finished chunks are formatted while the GPU runs the following chunks
                    for done, det, cnt, status in self.engine.infer_batch_stream([arrays[i] for i in good],
                                                                                detection_threshold):
                        for j in done:
                            finish(j, good[j], det, cnt, status)
            except Exception as e:
                print('Warning: batch inference failed: {}'.format(str(e)))
                for i in good:
                    if results[i] is None:
                        results[i] = {'file': ids[i], 'detections': None, 'failure': FAILURE_INFER}
        return results

    def generate_detections_from_jpeg_files(self, files, detection_threshold=0.00001, verbose=False):
        """
        JPEG files decoded on the GPU (nvJPEG) and detected in one call: the counterpart of `vis_utils.load_image` +
        `generate_detections_one_batch` (run_detector_batch.py:724,756) for `.jpg` files.  `files`: list of paths or of
        (image_id, bytes) pairs.  Same records; unreadable files get FAILURE_IMAGE_OPEN.  Decoded pixels differ from PIL's by
        a few LSB (different JPEG arithmetic), so confidences may differ in the third decimal.
        """
        if self.engine is None:
            raise RuntimeError('this B200Detector was created with preprocess_only; it cannot run inference')
        ids, blobs = [], []
        for f in files:
            if isinstance(f, (tuple, list)):
                ids.append(f[0])
                blobs.append(bytes(f[1]))
            else:
                ids.append(f)
                try:
                    with open(f, 'rb') as fh:
                        blobs.append(fh.read())
                except Exception:
                    blobs.append(None)
        if self._current_image_size != self.default_image_size:
            self.engine.set_image_size(self.default_image_size)
            self._current_image_size = self.default_image_size
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            det, cnt, _, status, hw = self.engine.infer_jpeg_batch(blobs, detection_threshold)
            results = []
            for i, iid in enumerate(ids):
                if status[i] == _capi.MDB_IMG_BAD_INPUT:
                    results.append({'file': iid, 'detections': None, 'failure': FAILURE_IMAGE_OPEN})
                elif status[i] != _capi.MDB_IMG_OK:
                    results.append({'file': iid, 'detections': None, 'failure': FAILURE_INFER})
                else:
                    dets, max_conf = format_detections(det[i], cnt[i], int(hw[i, 0]), int(hw[i, 1]),
                                                       self.use_model_native_classes, self.classic)
                    results.append({'file': iid, 'detections': dets, 'max_detection_conf': max_conf})
        finally:
            if gc_was_enabled:
                gc.enable()
        return results

    def generate_detections_one_image(self, img_original, image_id='unknown', detection_threshold=0.00001,
                                      image_size=None, augment=False, verbose=False):
        """pytorch_detector.py:1428-1478: wrapper around the batch function."""
        if isinstance(img_original, dict):
            batch_results = self.generate_detections_one_batch([img_original], None, detection_threshold, image_size,
                                                               augment, verbose)
        else:
            batch_results = self.generate_detections_one_batch([img_original], [image_id], detection_threshold,
                                                               image_size, augment, verbose)
        return batch_results[0]

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None
