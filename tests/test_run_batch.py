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
