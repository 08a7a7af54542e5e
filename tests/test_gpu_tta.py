"""
Test-time augmentation (`generate_detections_one_batch(..., augment=True)`, SURVEY.md par. 8f-4) against the oracle
restatement of yolov5's `_forward_augment` / `scale_img` / `_descale_pred` / `_clip_augmented` (oracle/yolov5.py; the package is
absent offline: parity unpinned like the rest of the forward oracle).  Reference call site:
/root/reference/megadetector/detection/pytorch_detector.py:1313 `self.model(batch_tensor, augment=augment)[0]`, followed by
the same nms() / scale_coords / formatting as without augmentation.
"""
import numpy as np
import pytest

from helpers import assert_anchors_equal_modulo_ties, assert_dets_equal_modulo_ties

pytestmark = pytest.mark.gpu


def _case(spec, arch, n_imgs, max_batch, sizes):
    import torch
    from megadetector_b200 import _capi, checkpoint
    from megadetector_b200.detector import B200Detector, format_detections
    from megadetector_b200.synth import synth_image
    from oracle import letterbox as olb
    from oracle import nms as onms
    from oracle import postprocess as opost
    from oracle import yolov5 as oy
    thr = 0.005
    det = B200Detector(spec, {'max_batch': max_batch})
    det.engine.set_pipeline_split(1)
    weights, _ = checkpoint.load_model(spec)
    imgs = [synth_image(i, *sizes[i % len(sizes)]) for i in range(n_imgs)]
    ids = ['t%d.jpg' % i for i in range(n_imgs)]
    plain = det.generate_detections_one_batch(imgs, ids, detection_threshold=thr)
    recs = det.generate_detections_one_batch(imgs, ids, detection_threshold=thr, augment=True)
    assert det.generate_detections_one_batch(imgs, ids, detection_threshold=thr) == plain  # the switch does not stick
    lbs = [olb.letterbox(im, 1280, 64, auto=True, scaleup=True)[0] for im in imgs]
    groups = {}
    for i, lb in enumerate(lbs):
        groups.setdefault(lb.shape, []).append(i)
    worst = 0.0
    for shape, idxs in groups.items():
        det.engine.set_augment(True)
        d, c, a, st = det.engine.infer_batch([imgs[i] for i in idxs], thr)
        pred = det.engine.fetch(_capi.MDB_FETCH_PRED)
        det.engine.set_augment(False)
        assert (st == 0).all()
        x = torch.from_numpy(np.stack([olb.to_model_input(lbs[i]) for i in idxs]))
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        wd = {k: v.cuda() for k, v in weights.items()}
        p_ref = oy.forward_augment(x.cuda(), wd, arch).cpu().numpy()
        assert pred.shape == p_ref.shape, (pred.shape, p_ref.shape)
        for j, i in enumerate(idxs):
            err = np.abs(pred[j] - p_ref[j])
            worst = max(worst, float(err[:, 4:].max()))
            assert float(err[:, 4:].max()) <= 1e-3, (i, float(err[:, 4:].max()))
            assert float(err[:, :4].max()) <= 1.28 / 0.67, (i, float(err[:, :4].max()))  # boxes are divided by the scale
            odet, oanc = onms.nms_torch_indexed(pred[j], thr, 0.45)
            assert c[j] == odet.shape[0]
            _, _, opx = opost.rescale_and_format(torch.from_numpy(odet), shape[:2], imgs[i].shape, thr)
            assert_dets_equal_modulo_ties(opx.numpy(), d[j, :c[j]], 'image %d' % i)
            assert_anchors_equal_modulo_ties(odet[:, 4], oanc, d[j, :c[j], 4], a[j, :c[j]], 'image %d' % i)
            gdets, gmax = format_detections(d[j], c[j], imgs[i].shape[0], imgs[i].shape[1])
            assert recs[i]['detections'] == gdets and recs[i]['max_detection_conf'] == gmax
    det.close()
    return worst


def test_tta_tiny_mixed_shapes():
    from oracle import yolov5 as oy
    w = _case('synthetic://tiny?seed=0&profile=contractive', oy.tiny_arch(), 5, 4, [(1080, 1920), (600, 800), (1080, 1920)])
    print('\n[tta tiny] max score err', w)


def test_tta_yolov5x6():
    from oracle import yolov5 as oy
    w = _case('synthetic://yolov5x6?seed=0&profile=contractive', oy.Arch(), 4, 4, [(1080, 1920)])
    print('\n[tta x6] max score err', w)
