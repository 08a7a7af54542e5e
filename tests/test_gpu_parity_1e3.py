"""
The north-star parity bar, asserted: on synthetic 1920x1080 inputs, at the benchmark batch
sizes bs=8 and bs=64, the engine's prediction tensor matches the fp32 oracle of the reference's PyTorch path
(/root/reference/megadetector/detection/pytorch_detector.py:1313) within
    max |d score| <= 1e-3      (objectness and class scores, all 61 200 x 4 values per image)
    max |d box|   <= 1e-3 * 1280 px   (xywh in letterbox pixels)
and the final result records agree detection by detection.

Weights: the CONTRACTIVE synthetic set (megadetector_b200/synth_weights.py, profile "contractive"; no MDv5 checkpoint is
available offline, and the forward oracle itself is a restatement: parity vs real MDv5a stays UNPINNED).  The round-1
"chaotic" set amplifies a last-bit perturbation ~6e4 times on its way to the heads -- no 10-bit-mantissa arithmetic, the
reference's own TF32 GPU default included, can be compared at 1e-3 on it; it stays in tests/test_gpu_e2e.py as the documented
stress case.  DESIGN.md par. 4 ("precision") states what the contractive set is and why each of its properties is needed.

What is compared, in this order:
  1. K1 output bit-exact (every image).
  2. prediction tensor vs the fp32 oracle: the two bars above, asserted.
  3. K6-K8 on the engine's own predictions: kept anchors / classes / boxes bit-exact vs the reference's nms() semantics.
  4. record level (md_tests.py:465-473,534-647 style, but matched by ANCHOR IDENTITY, which is stricter than IoU >= 0.85):
     every detection the oracle keeps with conf >= 0.05 must be kept by the engine with the same category,
     |d conf| <= CONF_TOL and |d coord| <= COORD_TOL -- unless the two candidates that compete for it are tied within the
     score tolerance in the oracle's own predictions (then which one survives NMS / the max_det cut is decided by the last
     bit in ANY implementation); such flips are counted, bounded and reported.
  5. "N of M images index-identical" is reported (and recorded in gpurun_out/parity_1e3_*.json).
"""
import json
import os

import numpy as np
import pytest

from helpers import assert_anchors_equal_modulo_ties, assert_dets_equal_modulo_ties

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-3            # BASELINE.json north_star
BOX_TOL_PX = 1e-3 * 1280    # 1e-3 of the letterbox size
CONF_TOL = 1.0e-3 + 1.0e-3 + 1e-9  # record confidences are truncated to 3 decimals (ct_utils.truncate_float): +1 unit of that
COORD_TOL = 2e-3                   # md_tests.py:100 is 1e-3; boxes are ROUNDED to whole pixels before they are normalised (:1379), one
                                   # pixel of a 1080-px side is 9.3e-4, and a width / height can move by two (both edges flip)
TIE_TOL = 2 * SCORE_TOL     # candidates closer than this in the ORACLE's own confidences may swap
MAX_FLIP_FRACTION = 0.03

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_pred(weights, arch, lbs, cpu_check=2):
    """fp32 oracle predictions.  The module graph is oracle.yolov5.forward in both cases; for speed the batch runs on the
    GPU through stock PyTorch with TF32 disabled (true fp32 multiply-accumulate), cross-checked against the CPU evaluation
    on the first images (they agree to ~1e-6, three orders below the bar)."""
    import torch
    from oracle import letterbox as olb
    from oracle import yolov5 as oy
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda')
    wd = {k: v.to(dev) for k, v in weights.items()}
    outs = []
    for i in range(0, len(lbs), 4):
        x = torch.from_numpy(np.stack([olb.to_model_input(lb) for lb in lbs[i:i + 4]])).to(dev)
        outs.append(oy.forward(x, wd, arch).cpu())
    pred = torch.cat(outs, 0).numpy()
    x = torch.from_numpy(np.stack([olb.to_model_input(lb) for lb in lbs[:cpu_check]]))
    cpu = oy.forward(x, {k: v.float() for k, v in weights.items()}, arch).numpy()
    dev_vs_cpu = float(np.abs(cpu - pred[:cpu_check])[:, :, 4:].max())
    assert dev_vs_cpu < 2e-5, dev_vs_cpu
    return pred, dev_vs_cpu


def _iou(a, b):
    iw = max(0.0, min(a[2], b[2]) - max(a[0], b[0]))
    ih = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = iw * ih
    u = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / u if u > 0 else 0.0


def _conf_cls_box(row):
    cls = int(np.argmax(row[5:] * row[4]))
    return float(row[5 + cls] * row[4]), cls, (row[0] - row[2] / 2, row[1] - row[3] / 2, row[0] + row[2] / 2, row[1] + row[3] / 2)


def _explain(anchor, p_ref, kept_here, kept_other, cut_conf, saturated):
    """
    Why is `anchor` kept on one side only?  Judged on the ORACLE's predictions.  Greedy NMS keeps a candidate unless a kept,
    overlapping, same-class candidate with a higher confidence exists; so a one-sided anchor has, on the other side, either
      'tie'      an overlapping kept competitor whose confidence is within TIE_TOL (order decided by the last bit), or whose
                 IoU sits on the 0.45 threshold, or the anchor itself sits on the max_det cut;
      'cascade'  an overlapping kept competitor that is itself one-sided (its own discrepancy is judged separately);
      None       nothing of the sort: a real disagreement.
    """
    conf, cls, box = _conf_cls_box(p_ref[anchor])
    if saturated and conf <= cut_conf + TIE_TOL:
        return 'tie'
    verdict = None
    for b in kept_other:
        conf2, cls2, box2 = _conf_cls_box(p_ref[b])
        if cls2 != cls:
            continue
        iou = _iou(box, box2)
        if iou <= 0.40:
            continue
        if abs(conf - conf2) <= TIE_TOL or abs(iou - 0.45) <= 0.02:
            return 'tie'
        if b not in kept_here:
            verdict = 'cascade'
    return verdict


def _parity_case(tag, model_spec, arch, n_imgs, max_batch):
    import torch
    from megadetector_b200 import _capi, checkpoint
    from megadetector_b200.detector import B200Detector, format_detections
    from megadetector_b200.synth import synth_image
    from oracle import letterbox as olb
    from oracle import nms as onms
    from oracle import postprocess as opost
    thr = 1e-5
    det = B200Detector(model_spec, {'max_batch': max_batch})
    det.engine.set_pipeline_split(1)  # one chunk, so that the debug fetches see the whole batch
    weights, _ = checkpoint.load_model(model_spec)
    imgs = [synth_image(i, 1080, 1920) for i in range(n_imgs)]
    ids = ['img%03d.jpg' % i for i in range(n_imgs)]
    d, c, a, st = det.engine.infer_batch(imgs, thr)
    assert (st == 0).all()
    stem = det.engine.fetch(_capi.MDB_FETCH_STEM_INPUT)
    pred = det.engine.fetch(_capi.MDB_FETCH_PRED)
    assert pred.shape[0] == n_imgs
    lbs = [olb.letterbox(im, 1280, 64, auto=True, scaleup=True)[0] for im in imgs]
    # 1. K1
    for i in range(n_imgs):
        assert np.array_equal(stem[i].view(np.uint16), olb.to_stem_layout_fp16(lbs[i]).view(np.uint16)), 'K1, image %d' % i
    # 2. predictions vs the fp32 oracle
    p_ref, dev_vs_cpu = _oracle_pred(weights, arch, lbs)
    err = np.abs(pred - p_ref)
    score_err = float(err[:, :, 4:].max())
    box_err = float(err[:, :, :4].max())
    report = {'case': tag, 'images': n_imgs, 'max_score_err': score_err, 'max_box_err_px': box_err,
              'mean_score_err': float(err[:, :, 4:].mean()), 'oracle_gpu_fp32_vs_cpu_fp32': dev_vs_cpu,
              'per_image_max_score_err': [float(err[i, :, 4:].max()) for i in range(n_imgs)]}
    # 3 + 4. NMS parity on the engine's own predictions; record-level comparison with the oracle
    records = det.generate_detections_one_batch(imgs, ids, detection_threshold=thr)
    identical, identical_hi, flips, n_ref, n_matched, n_extra, max_dconf, max_dcoord = 0, 0, {}, 0, 0, 0, 0.0, 0.0
    cands, unexplained = [], []
    for i in range(n_imgs):
        odet, oanc = onms.nms_torch_indexed(pred[i], thr, 0.45)  # reference semantics on the ENGINE's predictions
        assert c[i] == odet.shape[0], 'image %d' % i
        _, _, opx = opost.rescale_and_format(torch.from_numpy(odet), lbs[i].shape[:2], imgs[i].shape, thr)
        assert_dets_equal_modulo_ties(opx.numpy(), d[i, :c[i]], 'image %d' % i)
        assert_anchors_equal_modulo_ties(odet[:, 4], oanc, d[i, :c[i], 4], a[i, :c[i]], 'image %d' % i)
        cands.append(int((pred[i][:, 4] > thr).sum()))
        # the oracle's own result on ITS predictions
        rdet, ranc = onms.nms_torch_indexed(p_ref[i], thr, 0.45)
        rdets, _, rpx = opost.rescale_and_format(torch.from_numpy(rdet), lbs[i].shape[:2], imgs[i].shape, thr)
        gdets, _ = format_detections(d[i], c[i], 1080, 1920)
        assert gdets == records[i]['detections']
        e_anchor = {int(x): j for j, x in enumerate(a[i, :c[i]])}
        o_anchor = {int(x): j for j, x in enumerate(ranc)}
        if set(e_anchor) == set(o_anchor):
            identical += 1
        saturated = len(ranc) >= 300 or c[i] >= 300
        cut_conf = float(rdet[-1, 4]) if len(rdet) else 0.0
        # records are stored in ascending confidence order: row j of the kept list is record len-1-j
        hi_o, hi_e = set(), set()
        for anc, j in o_anchor.items():
            rec = rdets[len(rdets) - 1 - j]
            if rec['conf'] < 0.05:
                continue
            hi_o.add(anc)
            n_ref += 1
            if anc in e_anchor:
                got = gdets[len(gdets) - 1 - e_anchor[anc]]
                assert got['category'] == rec['category'], ('image %d anchor %d' % (i, anc), got, rec)
                n_matched += 1
                max_dconf = max(max_dconf, abs(got['conf'] - rec['conf']))
                max_dcoord = max(max_dcoord, max(abs(x - y) for x, y in zip(got['bbox'], rec['bbox'])))
            else:
                why = _explain(anc, p_ref[i], o_anchor, e_anchor, cut_conf, saturated)
                flips[why] = flips.get(why, 0) + 1
                if why is None:
                    unexplained.append((i, anc, 'oracle-only'))
        for anc, j in e_anchor.items():  # the other direction: detections only the engine reports
            got = gdets[len(gdets) - 1 - j]
            if got['conf'] < 0.05:
                continue
            hi_e.add(anc)
            if anc not in o_anchor:
                n_extra += 1
                why = _explain(anc, p_ref[i], e_anchor, o_anchor, cut_conf, saturated)
                if why is None:
                    unexplained.append((i, anc, 'engine-only'))
        # sets of anchors above the record threshold, ignoring members that sit within one truncation unit of 0.05
        if {x for x in hi_o if x in e_anchor or rdets[len(rdets) - 1 - o_anchor[x]]['conf'] > 0.052} == \
                {x for x in hi_e if x in o_anchor or gdets[len(gdets) - 1 - e_anchor[x]]['conf'] > 0.052}:
            identical_hi += 1
    n_flips = sum(flips.values())
    report.update({'images_index_identical_all_300': identical, 'images_index_identical_conf_ge_0.05': identical_hi,
                   'ref_dets_conf_ge_0.05': n_ref, 'matched_by_anchor': n_matched, 'engine_only_dets_conf_ge_0.05': n_extra,
                   'nms_flips': {str(k): v for k, v in flips.items()}, 'unexplained_flips': len(unexplained), 'max_dconf': max_dconf,
                   'max_dcoord': max_dcoord, 'candidates_per_image_min_max': [min(cands), max(cands)]})
    det.close()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'parity_1e3_{}.json'.format(tag)), 'w') as f:
        json.dump(report, f, indent=1)
    print('\n[parity 1e-3 {}] {}'.format(tag, {k: v for k, v in report.items() if k != 'per_image_max_score_err'}))
    print('[parity 1e-3 {}] {} of {} images index-identical over the detections with conf >= 0.05, {} of {} over all 300 '
          'kept rows'.format(tag, identical_hi, n_imgs, identical, n_imgs))
    assert score_err <= SCORE_TOL, report
    assert box_err <= BOX_TOL_PX, report
    assert n_ref > 0, 'the record-level comparison is vacuous: no oracle detection with conf >= 0.05'
    assert max_dconf <= CONF_TOL, report
    assert max_dcoord <= COORD_TOL, report
    assert not unexplained, (unexplained[:10], report)
    assert n_flips <= MAX_FLIP_FRACTION * n_ref, report
    return report


def test_yolov5x6_contractive_bs8():
    from oracle import yolov5 as oy
    _parity_case('x6_bs8', 'synthetic://yolov5x6?seed=0&profile=contractive', oy.Arch(), 8, 8)


def test_yolov5x6_contractive_bs64():
    from oracle import yolov5 as oy
    _parity_case('x6_bs64', 'synthetic://yolov5x6?seed=0&profile=contractive', oy.Arch(), 64, 64)


def test_mdv5b_seed_contractive_bs8():
    """BASELINE config 4: alternate weights (seed 1), same architecture."""
    from oracle import yolov5 as oy
    _parity_case('x6_seed1_bs8', 'synthetic://mdv5b?profile=contractive', oy.Arch(), 8, 8)


def test_tiny_contractive_mixed_sizes_bs5():
    from oracle import yolov5 as oy
    _parity_case('tiny_bs5', 'synthetic://tiny?seed=0&profile=contractive', oy.tiny_arch(), 5, 8)
