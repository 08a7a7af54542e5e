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
        B200Detector('synthetic://tiny', {'compatibility_mode': 'modern', 'preprocess_only': True})
    with pytest.raises(NotImplementedError):
        B200Detector('synthetic://tiny', {'force_cpu': True})
