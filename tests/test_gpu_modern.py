"""
The non-"classic" compatibility mode ("modern", SURVEY.md par. 8f-4) on the GPU, against the oracle restatement of
/root/reference/megadetector/detection/pytorch_detector.py:1036-1101 (INTER_AREA / INTER_LINEAR pre-resize, padding to
ceil(side / stride + 0.5) * stride), :1318-1321 (NMS IoU 0.6), :1372-1381 (scale_coords with ratio_pad) and :1395-1397
(round() instead of truncation).  The oracle's preprocessing is pinned to the real cv2 in tests/test_oracle_modern.py.
TTA (`augment=True`) is covered by tests/test_gpu_tta.py.
"""
import numpy as np
import pytest

from helpers import assert_anchors_equal_modulo_ties, assert_dets_equal_modulo_ties

pytestmark = pytest.mark.gpu

SHAPES = [(1080, 1920), (2160, 3840), (1440, 2560), (600, 800), (450, 600), (1000, 1500), (333, 517), (1280, 1280), (720, 1280),
          (1280, 720), (100, 300), (2448, 3264)]


@pytest.mark.parametrize('use_ceil', [False, True])
def test_k1_modern_bit_exact(use_ceil):
    from megadetector_b200 import _capi
    from megadetector_b200.synth import synth_image
    from oracle import letterbox as olb
    e = _capi.Engine(_capi.make_config(max_batch=2, compat_mode=1, use_ceil_for_resize=use_ceil))
    for i, (h, w) in enumerate(SHAPES):
        img = synth_image(i, h, w)
        got = e.op_letterbox(img)
        want = olb.to_stem_layout_fp16(olb.preprocess_modern(img, 1280, 64, use_ceil)['img_processed'])
        assert got.shape == want.shape, ((h, w), got.shape, want.shape)
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), 'K1 modern differs for {}x{}'.format(h, w)
    e.close()


def test_modern_mode_end_to_end():
    import torch
    from megadetector_b200 import _capi, checkpoint
    from megadetector_b200.detector import B200Detector, format_detections
    from megadetector_b200.synth import synth_image
    from oracle import letterbox as olb
    from oracle import nms as onms
    from oracle import postprocess as opost
    from oracle import yolov5 as oy
    spec = 'synthetic://tiny?seed=0&profile=contractive'
    det = B200Detector(spec, {'max_batch': 8, 'compatibility_mode': 'modern'})
    assert not det.classic and abs(det._cfg.iou_thres - 0.6) < 1e-6
    det.engine.set_pipeline_split(1)
    weights, _ = checkpoint.load_model(spec)
    sizes = [(1080, 1920), (1080, 1920), (2160, 3840), (450, 600), (600, 800)]
    imgs = [synth_image(i, *sizes[i]) for i in range(len(sizes))]
    ids = ['m%d.jpg' % i for i in range(len(imgs))]
    recs = det.generate_detections_one_batch(imgs, ids, detection_threshold=0.005)
    pre = [olb.preprocess_modern(im, 1280, 64) for im in imgs]
    groups = {}
    for i, p in enumerate(pre):
        groups.setdefault(p['img_processed'].shape, []).append(i)
    assert len(groups) == 2  # 768x1344 and 1024x1344
    for shape, idxs in groups.items():
        d, c, a, st = det.engine.infer_batch([imgs[i] for i in idxs], 0.005)
        assert (st == 0).all()
        stem = det.engine.fetch(_capi.MDB_FETCH_STEM_INPUT)
        pred = det.engine.fetch(_capi.MDB_FETCH_PRED)
        x = torch.from_numpy(np.stack([olb.to_model_input(pre[i]['img_processed']) for i in idxs]))
        p_ref = oy.forward(x, weights, oy.tiny_arch()).numpy()
        for j, i in enumerate(idxs):
            assert np.array_equal(stem[j].view(np.uint16), olb.to_stem_layout_fp16(pre[i]['img_processed']).view(np.uint16))
            err = np.abs(pred[j] - p_ref[j])
            assert float(err[:, 4:].max()) <= 1e-3 and float(err[:, :4].max()) <= 1.28, (i, float(err[:, 4:].max()))
            odet, oanc = onms.nms_torch_indexed(pred[j], 0.005, 0.6)
            assert c[j] == odet.shape[0]
            ratio_pad = ((pre[i]['img_original'].shape[0] / imgs[i].shape[0], pre[i]['img_original'].shape[1] / imgs[i].shape[1]),
                         pre[i]['letterbox_pad'])
            odets, omax, opx = opost.rescale_and_format(torch.from_numpy(odet), shape[:2], imgs[i].shape, 0.005, ratio_pad)
            assert_dets_equal_modulo_ties(opx.numpy(), d[j, :c[j]], 'image %d' % i)
            assert_anchors_equal_modulo_ties(odet[:, 4], oanc, d[j, :c[j], 4], a[j, :c[j]], 'image %d' % i)
            gdets, gmax = format_detections(d[j], c[j], imgs[i].shape[0], imgs[i].shape[1], classic=False)
            if len({x['conf'] for x in odets}) == len(odets):  # no confidence ties: order is defined
                assert gdets == odets, i
            assert recs[i]['detections'] == gdets and recs[i]['max_detection_conf'] == gmax
    # the producer path: host-side preprocess_image (cv2) -> dict -> engine, must give the same records
    pp = B200Detector(spec, {'preprocess_only': True, 'compatibility_mode': 'modern'})
    dicts = [pp.preprocess_image(im, iid) for im, iid in zip(imgs, ids)]
    for dd, p in zip(dicts, pre):
        assert np.array_equal(dd['img_processed'], p['img_processed']) and dd['letterbox_pad'] == p['letterbox_pad']
    assert det.generate_detections_one_batch(dicts, None, detection_threshold=0.005) == recs
    det.close()
