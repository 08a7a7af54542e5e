"""
ORACLE (test infrastructure only -- never imported by the product path).

Box rescaling and record formatting of the reference, restated op-for-op WITH TORCH so that tensor/scalar
type-promotion rules are the real ones:

  * `scale_coords` / `clip_coords` / `xyxy2xywh`: third-party yolov5 utils.general (ultralytics-yolov5 0.1.1, absent
    offline -> PARITY UNPINNED for these three, restated from the public source; call sites
    /root/reference/megadetector/detection/pytorch_detector.py:1379,1389)
  * the per-detection loop                       pytorch_detector.py:1384-1411
  * `truncate_float`, `convert_yolo_to_xywh`     /root/reference/megadetector/utils/ct_utils.py:82-103,255-270
    (pinned by the reference's own unit tests ct_utils.py:1299-1346,1460-1503, re-run in tests/test_oracle_post.py)
"""

import math

import torch

CONF_DIGITS = 3
COORD_DIGITS = 4


def truncate_float(x, precision=3):
    return math.floor(x * (10 ** precision)) / (10 ** precision)


def truncate_float_array(xs, precision=3):
    return [truncate_float(x, precision=precision) for x in xs]


def convert_yolo_to_xywh(yolo_box):
    x_center, y_center, width_of_box, height_of_box = yolo_box
    x_min = x_center - width_of_box / 2.0
    y_min = y_center - height_of_box / 2.0
    return [x_min, y_min, width_of_box, height_of_box]


def clip_coords(boxes, shape):
    boxes[:, 0].clamp_(0, shape[1])
    boxes[:, 1].clamp_(0, shape[0])
    boxes[:, 2].clamp_(0, shape[1])
    boxes[:, 3].clamp_(0, shape[0])


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None):
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    clip_coords(coords, img0_shape)
    return coords


def xyxy2xywh(x):
    y = x.clone()
    y[:, 0] = (x[:, 0] + x[:, 2]) / 2
    y[:, 1] = (x[:, 1] + x[:, 3]) / 2
    y[:, 2] = x[:, 2] - x[:, 0]
    y[:, 3] = x[:, 3] - x[:, 1]
    return y


def rescale_and_format(det, letterboxed_hw, scaling_shape, detection_threshold, ratio_pad=None):
    """
    det: torch float32 [N,6] in letterbox pixels (output of nms()).  Mirrors pytorch_detector.py:1352-1422.
    ratio_pad=None: "classic" (scale from the shapes, truncation); otherwise the non-classic branch (:1372-1381,1395-1397):
    ratio_pad = ((h_resized / h, w_resized / w), letterbox_pad), Python round() instead of truncation.
    Returns (detections list, max_conf, rounded det tensor).
    """
    detections = []
    max_conf = 0.0
    det = det.clone()
    classic = ratio_pad is None
    if len(det) > 0:
        gn = torch.tensor(scaling_shape)[[1, 0, 1, 0]]
        det[:, :4] = scale_coords(letterboxed_hw, det[:, :4], scaling_shape, ratio_pad).round()
        for *xyxy, conf, cls in reversed(det):
            if conf < detection_threshold:
                continue
            xywh = (xyxy2xywh(torch.tensor(xyxy).view(1, 4)) / gn).view(-1).tolist()
            api_box = convert_yolo_to_xywh(xywh)
            if classic:
                api_box = truncate_float_array(api_box, precision=COORD_DIGITS)
                conf = truncate_float(conf.tolist(), precision=CONF_DIGITS)
            else:
                api_box = [round(x, COORD_DIGITS) for x in api_box]
                conf = round(conf.tolist(), CONF_DIGITS)
            cls = int(cls.tolist()) + 1
            if cls not in (1, 2, 3):
                raise KeyError(f'{cls} is not a valid class.')
            detections.append({'category': str(cls), 'conf': conf, 'bbox': api_box})
            max_conf = max(max_conf, conf)
    return detections, max_conf, det
