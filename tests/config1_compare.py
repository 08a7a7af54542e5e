"""Comparison of two MegaDetector result files in the style of the reference's md_tests.compare_results
(/root/reference/megadetector/utils/md_tests.py:465-473,534-647): per file, detections are matched by category and
IoU >= 0.85; matched pairs are held to |d conf| <= 0.005 (md_tests.py:96) and |d coord| <= 0.002 (md_tests.py:100 is 0.001;
the reference rounds boxes to whole pixels before normalising them, pytorch_detector.py:1379, so one rounding flip per edge is
0.0017-0.0022 on these 450-600 px images).  Detections without a partner are counted: on the dense synthetic detector they
are NMS order flips between candidates that are tied within the score tolerance (tests/test_gpu_parity_1e3.py classifies
them on the prediction tensors); their share is bounded."""
import json

MAX_CONF_ERROR = 0.005
MAX_COORD_ERROR = 0.0025
MAX_UNMATCHED_FRACTION = 0.02
IOU_MATCH = 0.85


def _iou(a, b):
    ax2, ay2, bx2, by2 = a[0] + a[2], a[1] + a[3], b[0] + b[2], b[1] + b[3]
    iw = max(0.0, min(ax2, bx2) - max(a[0], b[0]))
    ih = max(0.0, min(ay2, by2) - max(a[1], b[1]))
    inter = iw * ih
    u = a[2] * a[3] + b[2] * b[3] - inter
    return inter / u if u > 0 else 0.0


def compare_files(actual_path, expected_path, conf_floor=0.05):
    a = json.load(open(actual_path))
    e = json.load(open(expected_path))
    assert a['detection_categories'] == e['detection_categories']
    fa = {im['file'].replace('\\', '/'): im for im in a['images']}
    fe = {im['file'].replace('\\', '/'): im for im in e['images']}
    assert sorted(fa) == sorted(fe), (sorted(fa), sorted(fe))
    stats = {'files': len(fa), 'n_expected': 0, 'matched': 0, 'unmatched': 0, 'max_conf_error': 0.0, 'max_coord_error': 0.0}
    for fn, ie in fe.items():
        ia = fa[fn]
        assert ('failure' in ia) == ('failure' in ie), fn
        if 'failure' in ie:
            continue
        pool = list(ia['detections'])
        for d in ie['detections']:
            if d['conf'] < conf_floor:
                continue
            stats['n_expected'] += 1
            best, bi = 0.0, None
            for j, g in enumerate(pool):
                if g['category'] != d['category']:
                    continue
                v = _iou(d['bbox'], g['bbox'])
                if v > best:
                    best, bi = v, j
            if bi is None or best < IOU_MATCH:
                stats['unmatched'] += 1
                continue
            g = pool.pop(bi)
            stats['matched'] += 1
            stats['max_conf_error'] = max(stats['max_conf_error'], abs(g['conf'] - d['conf']))
            stats['max_coord_error'] = max(stats['max_coord_error'], max(abs(x - y) for x, y in zip(g['bbox'], d['bbox'])))
    return stats


def assert_equivalent(actual_path, expected_path, what=''):
    s = compare_files(actual_path, expected_path)
    print('[config1 {}] {}'.format(what, s))
    assert s['n_expected'] > 0, s
    assert s['max_conf_error'] <= MAX_CONF_ERROR, s
    assert s['max_coord_error'] <= MAX_COORD_ERROR, s
    assert s['unmatched'] <= MAX_UNMATCHED_FRACTION * s['n_expected'] + 1, s
    return s
