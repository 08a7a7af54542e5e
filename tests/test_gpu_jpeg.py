"""
JPEG decode on the GPU (SURVEY.md par. 8f-1): nvJPEG -> K1, against the reference's host loader
`visualization_utils.load_image` (/root/reference/megadetector/visualization/visualization_utils.py:103-176,306-323: PIL /
libjpeg-turbo, RGB conversion, EXIF orientation 3/6/8), restated in megadetector_b200.run_batch.load_image.

nvJPEG and libjpeg-turbo use different IDCT and chroma up-sampling arithmetic, so pixels are compared with a tolerance (stated
below, measured values printed) and detections through the same record-level harness as BASELINE config 1.
"""
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAMPLES = [os.path.join(ROOT, 'tests', 'golden', 'config1_images', f)
           for f in ('nacti.jpg', 'idaho-camera-traps.jpg', 'orinoquia-thumb-web.jpg')]


@pytest.fixture(scope='module')
def engine():
    from megadetector_b200 import _capi
    e = _capi.Engine(_capi.make_config(max_batch=4))
    yield e
    e.close()


def _pil_rgb(blob):
    from PIL import Image
    from megadetector_b200 import run_batch
    return np.asarray(run_batch.load_image(io.BytesIO(blob)))


def _encode(arr, **kw):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(arr).save(b, format='JPEG', **kw)
    return b.getvalue()


def _with_orientation(blob, orientation):
    from PIL import Image
    im = Image.open(io.BytesIO(blob))
    exif = Image.Exif()
    exif[274] = orientation
    b = io.BytesIO()
    im.save(b, format='JPEG', exif=exif.tobytes(), quality=90)
    return b.getvalue()


def _cases():
    from megadetector_b200.synth import synth_image
    out = [(os.path.basename(p), open(p, 'rb').read()) for p in SAMPLES]
    img = synth_image(3, 1080, 1920)
    out.append(('synth_q90_420', _encode(img, quality=90)))
    out.append(('synth_q95_444', _encode(img, quality=95, subsampling=0)))
    out.append(('synth_q75_422', _encode(img, quality=75, subsampling=1)))
    out.append(('synth_progressive', _encode(img, quality=85, progressive=True)))
    out.append(('synth_grey', _encode(img[:, :, 1], quality=90)))
    out.append(('odd_size', _encode(synth_image(4, 333, 517), quality=90)))
    return out


def test_decoded_pixels_close_to_pil(engine):
    worst = {}
    for name, blob in _cases():
        got, backend = engine.op_decode_jpeg(blob)
        ref = _pil_rgb(blob)
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
        worst[name] = (backend, round(float(d.mean()), 3), int(d.max()), round(float((d > 2).mean()), 4), round(float((d > 16).mean()), 5))
    print('\n[nvjpeg vs PIL] name: (backend, mean |d|, max |d|, share > 2 LSB, share > 16 LSB)', worst)
    for name, w in worst.items():
        # different IDCT / chroma up-sampling arithmetic (libjpeg-turbo's "fancy" triangle filter vs nvJPEG's): small on
        # average, larger only at sharp chroma edges
        assert w[1] < 3.0, (name, w)
        assert w[4] < 0.01, (name, w)


@pytest.mark.parametrize('orientation', [1, 3, 6, 8, 2])
def test_exif_orientation_matches_pil(engine, orientation):
    from megadetector_b200.synth import synth_image
    blob = _with_orientation(_encode(synth_image(5, 240, 400), quality=92), orientation)
    got, _ = engine.op_decode_jpeg(blob)
    ref = _pil_rgb(blob)
    assert got.shape == ref.shape, (orientation, got.shape, ref.shape)
    d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    print('\n[nvjpeg orientation %d] mean |d| %.3f max %d' % (orientation, float(d.mean()), int(d.max())))
    assert d.mean() < 3.0 and (d > 16).mean() < 0.01, (orientation, float(d.mean()), int(d.max()))


def test_unreadable_and_cmyk_files_are_failures_not_exceptions():
    from PIL import Image
    from megadetector_b200.detector import B200Detector, FAILURE_IMAGE_OPEN
    from megadetector_b200.synth import synth_image
    det = B200Detector('synthetic://tiny?seed=0', {'max_batch': 4})
    good = _encode(synth_image(0, 270, 480), quality=90)
    b = io.BytesIO()
    Image.fromarray(synth_image(1, 64, 64)).convert('CMYK').save(b, format='JPEG')
    res = det.generate_detections_from_jpeg_files([('a.jpg', good), ('b.jpg', b'not a jpeg at all'), ('c.jpg', b.getvalue()),
                                                   ('d.jpg', good[:len(good) // 2]), '/nonexistent/e.jpg'])
    assert res[0]['detections'] is not None and [r['file'] for r in res] == ['a.jpg', 'b.jpg', 'c.jpg', 'd.jpg', '/nonexistent/e.jpg']
    for r in (res[1], res[2], res[4]):
        assert r['detections'] is None and r['failure'] == FAILURE_IMAGE_OPEN, r
    # a truncated stream is either decoded with the missing part grey (libjpeg behaviour) or reported unreadable
    assert res[3]['detections'] is not None or res[3]['failure'] == FAILURE_IMAGE_OPEN
    det.close()


def test_detections_from_gpu_decoded_files_match_the_pil_path():
    """Same record-level harness as BASELINE config 1: the three sample JPEGs + synthetic ones through nvJPEG -> K1 versus
    PIL -> K1, on the contractive yolov5x6 set."""
    import json
    from config1_compare import compare_files
    from megadetector_b200.detector import B200Detector
    from megadetector_b200 import run_batch
    det = B200Detector('synthetic://yolov5x6?seed=0&profile=contractive', {'max_batch': 8})
    cases = _cases()
    ids = [n for n, _ in cases]
    a = det.generate_detections_from_jpeg_files(cases)
    b = det.generate_detections_one_batch([run_batch.load_image(io.BytesIO(blob)) for _, blob in cases], ids)
    det.close()
    import tempfile
    tmp = tempfile.mkdtemp()
    for tag, recs in (('gpu', a), ('pil', b)):
        with open(os.path.join(tmp, tag + '.json'), 'w') as f:
            json.dump({'images': recs, 'detection_categories': {'1': 'animal', '2': 'person', '3': 'vehicle'}, 'info': {}}, f)
    s = compare_files(os.path.join(tmp, 'gpu.json'), os.path.join(tmp, 'pil.json'))
    print('\n[nvjpeg path vs PIL path, records]', s)
    assert s['n_expected'] > 100
    # a few LSB of pixel difference move confidences in the third decimal and flip near-tied NMS decisions
    # (matched pairs only need IoU >= 0.85: after an NMS flip the partner can be the neighbouring anchor)
    assert s['max_conf_error'] <= 0.01 and s['max_coord_error'] <= 0.05, s
    assert s['unmatched'] <= 0.1 * s['n_expected'], s
