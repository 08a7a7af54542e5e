"""CPU: host logic of the PTDetector mirror that needs no GPU: API-misuse errors, producer-side preprocessing."""
import numpy as np
import pytest

from megadetector_b200.detector import B200Detector, FAILURE_IMAGE_OPEN, FAILURE_INFER
from oracle import letterbox as olb


@pytest.fixture(scope='module')
def pre():
    return B200Detector('synthetic://tiny?seed=0', {'preprocess_only': True})


def test_constants_match_reference():
    # /root/reference/megadetector/detection/run_detector.py:55-67
    assert FAILURE_INFER == 'inference failure' and FAILURE_IMAGE_OPEN == 'image access failure'


def test_preprocess_only_instance_cannot_infer_and_has_reference_attributes(pre):
    assert pre.default_image_size == 1280 and pre.letterbox_stride == 64 and pre.compatibility_mode == 'classic'
    assert pre.device == 'cpu' and pre.use_model_native_classes is False
    with pytest.raises(RuntimeError):
        pre.generate_detections_one_batch([np.zeros((8, 8, 3), np.uint8)], ['a'])


@pytest.mark.parametrize('hw', [(1080, 1920), (450, 600), (600, 800)])
def test_preprocess_image_dict_matches_reference_contract(pre, hw):
    rng = np.random.default_rng(hw[0])
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    d = pre.preprocess_image(img, image_id='x.jpg')
    # keys of pytorch_detector.py:1111-1117
    assert set(d) == {'file', 'img_processed', 'img_original', 'img_original_pil', 'target_shape', 'scaling_shape',
                      'letterbox_ratio', 'letterbox_pad'}
    want, ratio, pad = olb.letterbox(img)
    assert np.array_equal(d['img_processed'], want)
    assert d['scaling_shape'] == img.shape and d['file'] == 'x.jpg' and d['img_original_pil'] is None
    assert tuple(d['letterbox_pad']) == tuple(float(p) for p in pad)


def test_api_misuse_raises_like_the_reference(pre):
    # pytorch_detector.py:1156-1182
    with pytest.raises(ValueError, match='must be a list'):
        pre.generate_detections_one_batch(np.zeros((4, 4, 3), np.uint8), ['a'])
    assert pre.generate_detections_one_batch([], []) == []
    with pytest.raises(ValueError, match='image_id must be a list'):
        pre.generate_detections_one_batch([np.zeros((4, 4, 3), np.uint8)], None)
    with pytest.raises(ValueError, match='Length mismatch'):
        pre.generate_detections_one_batch([np.zeros((4, 4, 3), np.uint8)], ['a', 'b'])
    with pytest.raises(ValueError, match='Mixed input types'):
        pre.generate_detections_one_batch([{'file': 'a'}, np.zeros((4, 4, 3), np.uint8)], None)
    with pytest.raises(ValueError, match='Mixed input types'):
        pre.generate_detections_one_batch([np.zeros((4, 4, 3), np.uint8), {'file': 'a'}], ['a', 'b'])


def test_unsupported_modes_fail_loudly():
    with pytest.raises(NotImplementedError):
        B200Detector('synthetic://tiny', {'force_cpu': True})


@pytest.mark.parametrize('hw', [(1080, 1920), (450, 600), (2160, 3840)])
def test_modern_mode_producer_preprocessing_matches_the_oracle(hw):
    """Non-classic compatibility mode on a producer worker (pytorch_detector.py:1036-1101): cv2 INTER_AREA / INTER_LINEAR
    pre-resize + pad, with the reference's dict keys (img_original = the PRE-RESIZED image, scaling_shape = the original)."""
    pre = B200Detector('synthetic://tiny?seed=0', {'preprocess_only': True, 'compatibility_mode': 'modern'})
    assert not pre.classic
    rng = np.random.default_rng(hw[1])
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    d = pre.preprocess_image(img, image_id='m.jpg')
    want = olb.preprocess_modern(img, 1280, 64)
    assert np.array_equal(d['img_processed'], want['img_processed'])
    assert np.array_equal(d['img_original'], want['img_original'])
    assert d['scaling_shape'] == img.shape and tuple(d['letterbox_pad']) == want['letterbox_pad']
    assert tuple(d['target_shape']) == tuple(want['target_shape'])


def test_format_detections_round_vs_truncate():
    """pytorch_detector.py:1392-1397: classic truncates (ct_utils.truncate_float), every other mode uses Python round()."""
    from megadetector_b200.detector import format_detections
    det = np.array([[100.0, 50.0, 301.0, 250.0, 0.98765, 0.0], [0.0, 0.0, 7.0, 9.0, 0.1235, 2.0]], np.float32)
    c, _ = format_detections(det, 2, 1080, 1920, classic=True)
    m, _ = format_detections(det, 2, 1080, 1920, classic=False)
    assert [x['conf'] for x in c] == [0.123, 0.987] and [x['conf'] for x in m] == [round(float(np.float32(0.1235)), 3), 0.988]
    x_min = (np.float32(100.0) + np.float32(301.0)) / np.float32(2) / np.float32(1920) - (np.float32(201.0) / np.float32(1920)) / 2.0
    assert c[1]['bbox'][0] == np.floor(float(x_min) * 1e4) / 1e4 and m[1]['bbox'][0] == round(float(x_min), 4)
    assert c[0]['category'] == '3' and m[1]['category'] == '1'


def test_c_formatter_equals_the_numpy_statement():
    """csrc/pyformat.c against format_detections_numpy (the restatement of pytorch_detector.py:1384-1411): every float32 /
    float64 step, truncation vs round(), key order, max conf, empty input, the KeyError of an unknown class."""
    from megadetector_b200.detector import format_detections, format_detections_numpy
    rng = np.random.default_rng(5)
    for trial in range(60):
        n = int(rng.integers(0, 301))
        hh, ww = int(rng.integers(100, 3000)), int(rng.integers(100, 4000))
        det = np.zeros((300, 6), np.float32)
        det[:, 0] = np.round(rng.uniform(0, ww, 300))
        det[:, 1] = np.round(rng.uniform(0, hh, 300))
        det[:, 2] = np.round(np.minimum(ww, det[:, 0] + rng.uniform(0, ww, 300)))
        det[:, 3] = np.round(np.minimum(hh, det[:, 1] + rng.uniform(0, hh, 300)))
        det[:, 4] = np.sort(rng.uniform(1e-5, 1, 300).astype(np.float32))[::-1]
        det[:, 5] = rng.integers(0, 3, 300)
        for classic in (True, False):
            for native in (False, True):
                got = format_detections(det, n, hh, ww, native, classic)
                want = format_detections_numpy(det, n, hh, ww, native, classic)
                assert got == want
                assert [list(d) for d in got[0]] == [['category', 'conf', 'bbox']] * n
    det[7, 5] = 5.0
    for f in (format_detections, format_detections_numpy):
        with pytest.raises(KeyError, match='6 is not a valid class'):
            f(det, 300, 1080, 1920)
        assert f(det, 300, 1080, 1920, True)[0][300 - 1 - 7]['category'] == '5'
        assert f(det, 0, 1080, 1920) == ([], 0.0)
    with pytest.raises(ValueError):
        from megadetector_b200 import detector
        detector._formatter().format_detections(np.zeros((2, 6), np.float32), 3, 10, 10, False, True)
