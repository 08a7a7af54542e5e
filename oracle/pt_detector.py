"""
ORACLE (test infrastructure only -- never imported by the product path).

`OraclePTDetector`: the CPU oracle of the hot path wearing the reference's detector duck type
(`PTDetector`, /root/reference/megadetector/detection/pytorch_detector.py:739-1482: constructor signature,
`preprocess_image` :964, `generate_detections_one_batch` :1124, `generate_detections_one_image` :1428), so that the
reference's OWN drivers (`run_detector_batch.load_and_run_detector_batch` :1062, the image-queue producers :124-200) can run
the CPU arm of BASELINE.json's config 1 end to end.  The real `PTDetector` cannot: its module graph lives in the un-vendored
`ultralytics-yolov5` package (parity UNPINNED, see oracle/yolov5.py).
"""
import numpy as np
import torch

from megadetector_b200 import checkpoint

from . import letterbox as olb
from . import nms as onms
from . import postprocess as opost
from . import yolov5 as oyolo
from .detector import FAILURE_IMAGE_OPEN, FAILURE_INFER, OracleDetector


class OraclePTDetector:
    def __init__(self, model_path, detector_options=None, verbose=False):
        opts = detector_options or {}
        self.preprocess_only = bool(opts.get('preprocess_only'))
        self.default_image_size = 1280
        self.letterbox_stride = 64
        self.ora = None
        if not self.preprocess_only:
            w, arch = checkpoint.load_model(model_path)
            a = oyolo.Arch(ch=list(arch['ch']), n_backbone=list(arch['n_backbone']), n_head=arch['n_head'], nc=arch['nc'])
            if arch.get('anchors_px'):
                a.anchors_px = [list(x) for x in arch['anchors_px']]
            self.ora = OracleDetector(w, a, fast_nms=True)
            print('PTDetector using device cpu')

    def preprocess_image(self, img_original, image_id='unknown', image_size=None, verbose=False):
        """pytorch_detector.py:964-1119 ("classic")."""
        img_original_pil = None
        if not isinstance(img_original, np.ndarray):
            img_original_pil = img_original
            img_original = np.asarray(img_original)
        size = self.default_image_size if image_size is None else image_size
        lb, ratio, pad = olb.letterbox(img_original, size, self.letterbox_stride, auto=True, scaleup=True)
        return {'file': image_id, 'img_processed': lb, 'img_original': img_original, 'img_original_pil': img_original_pil,
                'target_shape': size, 'scaling_shape': img_original.shape, 'letterbox_ratio': ratio, 'letterbox_pad': pad}

    def generate_detections_one_batch(self, img_original, image_id=None, detection_threshold=0.00001, image_size=None,
                                      augment=False, verbose=False):
        if not isinstance(img_original, list):
            raise ValueError('img_original must be a list for batch processing')
        if len(img_original) == 0:
            return []
        if isinstance(img_original[0], dict):
            pre = img_original
        else:
            pre = []
            for im, iid in zip(img_original, image_id):
                try:
                    a = np.asarray(im)
                    if a.ndim != 3 or a.shape[2] != 3:
                        raise ValueError('expected an HxWx3 image')
                    pre.append(self.preprocess_image(a, iid, image_size))
                except Exception:
                    pre.append({'file': iid, 'failure': FAILURE_IMAGE_OPEN})
        results = [None] * len(pre)
        groups = {}
        for i, p in enumerate(pre):
            if 'failure' in p:
                results[i] = {'file': p['file'], 'detections': None, 'failure': p['failure']}
            else:
                groups.setdefault(p['img_processed'].shape, []).append(i)
        for shape, idxs in groups.items():
            try:
                pred = self.ora.forward_letterboxed([pre[i]['img_processed'] for i in idxs])
                for j, i in enumerate(idxs):
                    det = onms.nms_torch(torch.from_numpy(pred[j:j + 1]), detection_threshold, 0.45)[0]
                    dets, max_conf, _ = opost.rescale_and_format(det, shape[:2], pre[i]['scaling_shape'], detection_threshold)
                    results[i] = {'file': pre[i]['file'], 'detections': dets, 'max_detection_conf': max_conf}
            except Exception as e:
                print('Warning: batch inference failed: {}'.format(e))
                for i in idxs:
                    results[i] = {'file': pre[i]['file'], 'detections': None, 'failure': FAILURE_INFER}
        return results

    def generate_detections_one_image(self, img_original, image_id='unknown', detection_threshold=0.00001, image_size=None,
                                      augment=False, verbose=False):
        if isinstance(img_original, dict):
            return self.generate_detections_one_batch([img_original], None, detection_threshold, image_size)[0]
        return self.generate_detections_one_batch([img_original], [image_id], detection_threshold, image_size)[0]
