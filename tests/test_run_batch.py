"""
The in-process multi-GPU driver (megadetector_b200/run_batch.py; SURVEY.md par. 8f-3, the reference author's TODO at
/root/reference/TODO.md:203-205).  CPU: the plumbing (chunking, spawn workers, merge, uniqueness, result file layout, failure
records) with a stand-in detector.  GPU (2 GPUs): the records of a 2-GPU run equal those of a 1-GPU run.
"""
import json
import os

import numpy as np
import pytest

from megadetector_b200 import run_batch, sharding


class FakeDetector:
    """Stand-in with the detector duck type: one deterministic detection per image, conf from the image content."""

    def __init__(self, model_file, options):
        self.device = options['device']

    def generate_detections_one_batch(self, images, ids, detection_threshold=1e-5):
        out = []
        for im, f in zip(images, ids):
            a = np.asarray(im)
            conf = float(int(a.sum()) % 1000) / 1000.0
            out.append({'file': f, 'max_detection_conf': conf,
                        'detections': [{'category': '1', 'conf': conf, 'bbox': [0.1, 0.2, 0.3, 0.4]},
                                       {'category': '2', 'conf': 0.001, 'bbox': [0, 0, 1, 1]}]})
        return out


def _make_images(folder, n):
    from PIL import Image
    rng = np.random.default_rng(0)
    files = []
    for i in range(n):
        sub = os.path.join(folder, 'cam%d' % (i % 3))
        os.makedirs(sub, exist_ok=True)
        p = os.path.join(sub, 'img%03d.jpg' % i)
        Image.fromarray(rng.integers(0, 255, size=(40 + i, 60, 3), dtype=np.uint8)).save(p)
        files.append(p)
    with open(os.path.join(folder, 'cam0', 'broken.jpg'), 'w') as f:
        f.write('not a jpeg')
    return files


def test_sharded_driver_plumbing_on_cpu(tmp_path):
    files = _make_images(str(tmp_path), 11)
    found = run_batch.find_images(str(tmp_path), recursive=True)
    assert len(found) == 12 and found == sorted(found)
    one, _ = run_batch.run_sharded('m.pt', found, 1, batch_size=4, detector_factory='test_run_batch:FakeDetector')
    three, times = run_batch.run_sharded('m.pt', found, 3, batch_size=4, detector_factory='test_run_batch:FakeDetector')
    assert one == three and len(times) == 3
    assert [r['file'] for r in three] == found
    bad = [r for r in three if 'failure' in r]
    assert len(bad) == 1 and bad[0]['file'].endswith('broken.jpg') and bad[0]['failure'] == 'image access failure'
    # the output threshold is applied (run_detector_batch.py:766-767)
    assert all(len(r['detections']) <= 1 for r in three if 'failure' not in r)
    out = run_batch.write_results(three, str(tmp_path / 'out.json'), relative_path_base=str(tmp_path), detector_file='x/m.pt')
    d = json.load(open(out))
    assert set(d) == {'images', 'detection_categories', 'info'} and d['info']['detector'] == 'm.pt'
    assert d['images'][0]['file'] == 'cam0/broken.jpg' and 'max_detection_conf' not in d['images'][1]
    # chunking rule == ct_utils.split_list_into_n_chunks('greedy')
    assert [len(c) for c in sharding.split_list_into_n_chunks(found, 5)] == [3, 3, 2, 2, 2]


@pytest.mark.gpu
def test_two_gpus_give_the_single_gpu_records(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    from PIL import Image
    from megadetector_b200.synth import synth_image
    files = []
    for i in range(10):
        p = str(tmp_path / ('img%02d.png' % i))
        Image.fromarray(synth_image(i, 270 + 30 * (i % 3), 480)).save(p)
        files.append(p)
    spec = 'synthetic://tiny?seed=0'
    one, _ = run_batch.run_sharded(spec, files, 1, batch_size=4)
    two, _ = run_batch.run_sharded(spec, files, 2, batch_size=4)
    assert one == two and len(two) == 10 and all('detections' in r for r in two)


class FlakyDetector(FakeDetector):
    """Dies on the batch that contains the file named in MDB_TEST_FAIL_ON (a crash in the middle of a job)."""

    def generate_detections_one_batch(self, images, ids, detection_threshold=1e-5):
        bad = os.environ.get('MDB_TEST_FAIL_ON', '')
        if bad and any(f.endswith(bad) for f in ids):
            raise RuntimeError('simulated crash')
        return super().generate_detections_one_batch(images, ids, detection_threshold)


def test_checkpoint_and_resume_use_the_reference_format(tmp_path, monkeypatch):
    """run_detector_batch.py:1465-1520: {"checkpoint": [records]}; a job that dies resumes from it and ends with the records of
    an uninterrupted run; nothing is processed twice."""
    _make_images(str(tmp_path / 'in'), 11)
    found = [os.path.abspath(f) for f in run_batch.find_images(str(tmp_path / 'in'))]
    want, _ = run_batch.run_sharded('m.pt', found, 1, batch_size=2, detector_factory='test_run_batch:FakeDetector')
    ckpt = str(tmp_path / 'md_checkpoint_test.json')
    monkeypatch.setenv('MDB_TEST_FAIL_ON', 'img008.jpg')  # the last batch of worker 1
    with pytest.raises(RuntimeError, match='simulated crash'):
        run_batch.run_sharded('m.pt', found, 2, batch_size=2, detector_factory='test_run_batch:FlakyDetector',
                              checkpoint_path=ckpt, checkpoint_frequency=3)
    assert not os.path.exists(ckpt + '_tmp')
    data = json.load(open(ckpt))
    assert set(data) == {'checkpoint'} and 0 < len(data['checkpoint']) < len(found)
    assert all(os.path.isabs(r['file']) for r in data['checkpoint'])
    monkeypatch.delenv('MDB_TEST_FAIL_ON')
    restored = run_batch.load_checkpoint(ckpt)
    got, _ = run_batch.run_sharded('m.pt', found, 2, batch_size=2, detector_factory='test_run_batch:FlakyDetector',
                                   checkpoint_path=ckpt, checkpoint_frequency=3, results=restored)
    assert got == want
    # everything already in the checkpoint: no worker is started
    again, times = run_batch.run_sharded('m.pt', found, 2, batch_size=2, detector_factory='no_such_module:X', results=got)
    assert again == want and times == []
    with pytest.raises(ValueError, match='missing "checkpoint" field'):
        bad = tmp_path / 'bad.json'
        bad.write_text('{"images": []}')
        run_batch.load_checkpoint(str(bad))


def test_checkpoints_interoperate_with_the_reference_driver(tmp_path):
    ref_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'megadetector')):
        pytest.skip('baseline/_ref is not installed (tools/install_reference.sh)')
    import sys
    for p in (ref_root, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from megadetector.detection import run_detector_batch as ref
    recs = [{'file': '/a/b.jpg', 'detections': [{'category': '1', 'conf': 0.5, 'bbox': [0.1, 0.2, 0.3, 0.4]}]},
            {'file': '/a/c.jpg', 'failure': 'image access failure'}]
    ours, theirs = str(tmp_path / 'ours.json'), str(tmp_path / 'theirs.json')
    run_batch.write_checkpoint(ours, recs)
    ref.write_checkpoint(theirs, recs)
    assert ref.load_checkpoint(ours) == recs and run_batch.load_checkpoint(theirs) == recs
    assert json.load(open(ours)) == json.load(open(theirs))


def test_cli_checkpoint_flags(tmp_path, monkeypatch, capsys):
    """--checkpoint_frequency / --checkpoint_path / --resume_from_checkpoint auto / overwrite guard / deletion at the end
    (run_detector_batch.py:1983-2003, 2095-2118, 2179-2181)."""
    _make_images(str(tmp_path / 'in'), 7)
    out_dir = tmp_path / 'out'
    out_dir.mkdir()
    real = run_batch.run_sharded
    monkeypatch.setattr(run_batch, 'run_sharded',
                        lambda *a, **k: real(*a, detector_factory='test_run_batch:FakeDetector', **k))
    found = [os.path.abspath(f) for f in run_batch.find_images(str(tmp_path / 'in'))]
    first, _ = real('m.pt', found[:3], 1, batch_size=2, detector_factory='test_run_batch:FakeDetector')
    run_batch.write_checkpoint(str(out_dir / 'md_checkpoint_20250101000000.json'), first)
    with pytest.raises(AssertionError, match='already exists'):
        run_batch.main(['m.pt', str(tmp_path / 'in'), str(out_dir / 'o.json'), '--recursive', '--checkpoint_frequency', '2',
                        '--checkpoint_path', str(out_dir / 'md_checkpoint_20250101000000.json')])
    rc = run_batch.main(['m.pt', str(tmp_path / 'in'), str(out_dir / 'o.json'), '--recursive', '--batch_size', '2',
                         '--checkpoint_frequency', '2', '--resume_from_checkpoint', 'auto',
                         '--checkpoint_path', str(out_dir / 'ck.json'), '--output_relative_filenames'])
    assert rc == 0
    text = capsys.readouterr().out
    assert 'Restored 3 entries' in text and 'Bypassing 3 images' in text and 'Deleted checkpoint file' in text
    assert not (out_dir / 'ck.json').exists()
    d = json.load(open(out_dir / 'o.json'))
    assert len(d['images']) == 8 and d['images'][0]['file'] == 'cam0/broken.jpg'
    with pytest.raises(ValueError, match='no checkpoints found'):
        run_batch.find_checkpoint(str(tmp_path / 'in'))


def test_image_file_argument_forms(tmp_path):
    files = _make_images(str(tmp_path), 4)
    (tmp_path / 'l.json').write_text(json.dumps(files[:3]))
    (tmp_path / 'l.txt').write_text('\n'.join(files[1:]) + '\n\n')
    assert run_batch.resolve_image_list(str(tmp_path / 'l.json')) == files[:3]
    assert run_batch.resolve_image_list(str(tmp_path / 'l.txt')) == files[1:]
    assert run_batch.resolve_image_list(files[2]) == [files[2]]
    assert len(run_batch.resolve_image_list(str(tmp_path))) == 5
    with pytest.raises(ValueError):
        run_batch.resolve_image_list(str(tmp_path / 'nothing.bin'))


def test_results_file_equals_the_reference_writer(tmp_path):
    """write_results against the reference's own write_results_to_file (run_detector_batch.py:1546-1662) on the same records:
    byte-identical files except for the completion time."""
    ref_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'megadetector')):
        pytest.skip('baseline/_ref is not installed (tools/install_reference.sh)')
    import copy
    import re
    import sys
    for p in (ref_root, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from megadetector.detection import run_detector_batch as ref
    base = str(tmp_path)
    recs = [{'file': os.path.join(base, 'b', 'z.jpg'), 'max_detection_conf': 0.9,
             'detections': [{'category': '2', 'conf': 0.1, 'bbox': [0.1, 0.2, 0.3, 0.4]},
                            {'category': '1', 'conf': 0.9, 'bbox': [0.5, 0.5, 0.25, 0.125]},
                            {'category': '3', 'conf': 0.1, 'bbox': [0.0, 0.0, 1.0, 1.0]}]},
            {'file': os.path.join(base, 'a.jpg'), 'failure': 'image access failure'},
            {'file': os.path.join(base, 'b', 'c.jpg'), 'max_detection_conf': 0.0, 'detections': []}]
    for det_file in ('/models/md_v5a.0.1.pt', '/models/MDV5B.pt', 'synthetic://yolov5x6?seed=0', None):
        for rel in (None, base):
            ours, theirs = str(tmp_path / 'ours.json'), str(tmp_path / 'theirs.json')
            run_batch.write_results(copy.deepcopy(recs), ours, relative_path_base=rel, detector_file=det_file)
            ref.write_results_to_file(copy.deepcopy(recs), theirs, relative_path_base=rel, detector_file=det_file)
            strip = lambda t: re.sub(r'"detection_completion_time": "[^"]*"', '"detection_completion_time": ""', t)
            assert strip(open(ours).read()) == strip(open(theirs).read()), (det_file, rel)
