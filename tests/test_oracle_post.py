"""CPU: record formatting -- reference known answers, and the product's vectorised formatter vs the oracle loop."""
import numpy as np
import torch

from oracle import postprocess as opost


def test_reference_known_answers_for_truncation_and_box_conversion():
    # /root/reference/megadetector/utils/ct_utils.py:1299-1346 and :1460-1503
    assert opost.truncate_float(0.12345, precision=3) == 0.123
    assert opost.truncate_float(1.999, precision=2) == 1.99
    assert opost.truncate_float(0.0003214884) == 0.0
    assert opost.truncate_float(1.0003214884, precision=6) == 1.000321
    b = opost.convert_yolo_to_xywh([0.5, 0.5, 0.2, 0.2])
    assert np.allclose(b, [0.4, 0.4, 0.2, 0.2])


def _random_dets(rng, n, lb_hw):
    x1 = rng.uniform(-20, lb_hw[1], n)
    y1 = rng.uniform(-20, lb_hw[0], n)
    w = rng.uniform(0, 500, n)
    h = rng.uniform(0, 500, n)
    conf = np.sort(rng.uniform(1e-5, 1, n))[::-1]
    cls = rng.integers(0, 3, n)
    return np.stack([x1, y1, x1 + w, y1 + h, conf, cls], 1).astype(np.float32)


def test_vectorised_formatter_equals_reference_loop():
    from megadetector_b200.detector import format_detections
    rng = np.random.default_rng(3)
    for (src, lb) in [((1080, 1920), (768, 1280)), ((450, 600), (960, 1280)), ((600, 800), (960, 1280)),
                      ((1281, 1919), (896, 1280)), ((3000, 4000), (960, 1280))]:
        det = _random_dets(rng, 300, lb)
        dets_o, max_o, det_px = opost.rescale_and_format(torch.from_numpy(det), lb, (src[0], src[1], 3), 1e-5)
        dets_p, max_p = format_detections(det_px.numpy(), det_px.shape[0], src[0], src[1])
        assert dets_o == dets_p  # exact: same doubles, same strings, same order
        assert max_o == max_p
        assert all(isinstance(d['conf'], float) for d in dets_p)


def test_scale_coords_geometry_for_benchmark_shape():
    # 1920x1080 in 768x1280: gain 2/3, pad (0, 24)
    det = torch.tensor([[0.0, 24.0, 1280.0, 744.0, 0.9, 0.0], [640.2, 384.7, 700.9, 500.1, 0.5, 1.0]])
    _, _, px = opost.rescale_and_format(det, (768, 1280), (1080, 1920, 3), 1e-5)
    assert px[0, :4].tolist() == [0.0, 0.0, 1920.0, 1080.0]
    assert px[1, :4].tolist() == [960.0, 541.0, 1051.0, 714.0]
