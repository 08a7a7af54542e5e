"""CPU: the N>1 host logic (image sharding, result gather, max-over-ranks) with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from megadetector_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_detect(name):
    # deterministic stand-in for one image's record (the GPU engine is not needed to test the plumbing)
    h = sum(ord(c) for c in name)
    return {'file': name, 'max_detection_conf': (h % 1000) / 1000.0,
            'detections': [{'category': str(1 + h % 3), 'conf': (h % 1000) / 1000.0, 'bbox': [0.1, 0.2, 0.3, 0.4]}]}


def _worker(rank, world, port, files, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = sharding.shard_for_rank(files, rank, world)
    local = [_fake_detect(f) for f in mine]
    merged = sharding.gather_results(local, rank, world)
    tmax = sharding.max_over_ranks([10.0 + rank, 5.0 - rank], world)
    assert tmax == [10.0 + world - 1, 5.0]
    if rank == 0:
        torch.save({'merged': merged, 'n_local': len(local)}, out_path)
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_merge(tmp_path):
    files = ['img%03d.jpg' % i for i in range(37)][::-1]
    out = str(tmp_path / 'merged.pt')
    mp.spawn(_worker, args=(2, _free_port(), files, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    want = sorted((_fake_detect(f) for f in files), key=lambda r: r['file'])
    assert got['merged'] == want
    assert got['n_local'] == 19  # 37 = 19 + 18


def test_chunking_matches_reference_semantics():
    # ct_utils.split_list_into_n_chunks(greedy): contiguous, first chunks full
    assert sharding.split_list_into_n_chunks(list(range(10)), 3) == [[0, 1, 2, 3], [4, 5, 6], [7, 8, 9]]
    assert sharding.split_list_into_n_chunks([], 4) == [[], [], [], []]
    assert sharding.shard_for_rank(list(range(5)), 7, 8) == [] and sharding.shard_for_rank(list(range(5)), 4, 8) == [4]
