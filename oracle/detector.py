"""
ORACLE (test infrastructure only -- never imported by the product path).

The reference's whole hot path on the CPU, fp32, assembled from the restated pieces in the order
`PTDetector.generate_detections_one_batch` runs them
(/root/reference/megadetector/detection/pytorch_detector.py:1124-1424):
letterbox (:1104) -> group by shape (:1226) -> /255 NCHW fp32 (:1286-1306) -> model (:1313) -> nms (:1342)
-> scale_coords().round() (:1379) -> per-detection formatting (:1384-1411).

Used by the parity tests as the checker and by bench.py as the timed CPU baseline ("port").
"""

import numpy as np
import torch

from . import letterbox as olb
from . import nms as onms
from . import postprocess as opost
from . import yolov5 as oyolo

FAILURE_INFER = 'inference failure'
FAILURE_IMAGE_OPEN = 'image access failure'


class OracleDetector:
    def __init__(self, weights, arch=None, img_size=1280, stride=64, fast_nms=False):
        self.w = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))).float() for k, v in weights.items()}
        self.arch = arch or oyolo.Arch()
        self.img_size = img_size
        self.stride = stride
        self.last_pred = None
        self.fast_nms = fast_nms  # timed baseline: torchvision C++ NMS, as the reference uses

    def forward_letterboxed(self, imgs_u8):
        """list of equally-shaped letterboxed HWC u8 -> pred float32 numpy [B, anchors, 5+nc]"""
        x = torch.from_numpy(np.stack([olb.to_model_input(i) for i in imgs_u8]))
        return oyolo.forward(x, self.w, self.arch).numpy()

    def generate_detections_one_batch(self, images, image_ids, detection_threshold=1e-5, return_raw=False):
        n = len(images)
        results = [None] * n
        raw = [None] * n
        groups = {}
        for i, img in enumerate(images):
            try:
                lb, ratio, pad = olb.letterbox(np.asarray(img), self.img_size, self.stride, auto=True, scaleup=True)
                groups.setdefault(lb.shape, []).append((i, lb))
            except Exception:
                results[i] = {'file': image_ids[i], 'detections': None, 'failure': FAILURE_IMAGE_OPEN}
        for shape, items in groups.items():
            pred = self.forward_letterboxed([lb for _, lb in items])
            self.last_pred = pred
            for j, (i, lb) in enumerate(items):
                if self.fast_nms:
                    det = onms.nms_torch(torch.from_numpy(pred[j:j + 1]), detection_threshold, 0.45)[0].numpy()
                    anc = None
                else:
                    det, anc = onms.nms_one_image(pred[j], detection_threshold, 0.45)
                dets, max_conf, det_px = opost.rescale_and_format(torch.from_numpy(det), shape[:2],
                                                                  np.asarray(images[i]).shape, detection_threshold)
                results[i] = {'file': image_ids[i], 'detections': dets, 'max_detection_conf': max_conf}
                raw[i] = {'det_px': det_px.numpy(), 'anchors': anc, 'pred': pred[j], 'letterboxed': lb}
        return (results, raw) if return_raw else results
