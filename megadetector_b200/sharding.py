"""
Image sharding across ranks (one process per GPU) and the host-side result merge.

The reference scales out by splitting the file list into contiguous chunks, running one CLI process per GPU
(CUDA_VISIBLE_DEVICES=i) and concatenating the JSON results with uniqueness checks
(/root/reference/megadetector/detection/run_detector_batch.py:24-32,
 /root/reference/notebooks/manage_local_batch.py:496-547,933-975; chunking = ct_utils.split_list_into_n_chunks,
 utils/ct_utils.py:478-506, "greedy" contiguous).  Here the same thing happens inside a torch.distributed job:
every rank owns one mdb_engine, images are independent, and there is NO collective on the data path -- only the
final gather of result records to rank 0 (tiny: <= 300 boxes per image) and, in bench.py, a barrier and a
max-reduction of the timings.
"""

def split_list_into_n_chunks(items, n):
    """
    Contiguous chunks whose sizes differ by at most one, the first (len % n) chunks being the larger ones --
    the semantics of ct_utils.split_list_into_n_chunks(chunk_strategy='greedy').
    """
    n = max(1, int(n))
    base, extra = divmod(len(items), n)
    out, start = [], 0
    for i in range(n):
        size = base + (1 if i < extra else 0)
        out.append(items[start:start + size])
        start += size
    return out


def shard_for_rank(items, rank, world):
    """The contiguous chunk of `items` that rank `rank` of `world` processes."""
    return split_list_into_n_chunks(list(items), world)[rank]


def gather_results(local_results, rank, world, group=None):
    """
    Collect per-rank lists of result records on rank 0 (other ranks get None), check that no image was processed
    twice (manage_local_batch.py:947-958) and sort by 'file' (write_results_to_file does the same,
    run_detector_batch.py:1631).
    """
    if world == 1:
        merged = list(local_results)
    else:
        import torch.distributed as dist
        buckets = [None] * world if rank == 0 else None
        dist.gather_object(local_results, buckets, dst=0, group=group)
        if rank != 0:
            return None
        merged = [r for b in buckets for r in b]
    files = [r['file'] for r in merged]
    if len(set(files)) != len(files):
        raise ValueError('duplicate images in the merged results')
    return sorted(merged, key=lambda r: r['file'])


def max_over_ranks(values, world, device=None, group=None):
    """Element-wise maximum of a list of floats over all ranks (device timing convention of bench.py)."""
    if world == 1:
        return [float(v) for v in values]
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(v) for v in t.tolist()]
