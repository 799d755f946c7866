"""Multi-GPU plumbing: one process per GPU, graph replicated in every GPU's HBM, start nodes and
SimRank queries sharded in contiguous ranges, results gathered with torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  There is NO data-path collective: walks and
queries are independent given the read-only graph (node2vec.py:53-57, SingleRandomWalk.java:39-45);
the only exchange is the final gather of fixed-size result tiles.

Random streams are keyed by GLOBAL walk / query index (Philox counter), so the gathered result is
identical for any world size.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n units for this rank; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def _dist():
    import torch.distributed as dist
    return dist


def world_info():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_rows(local, n_total, device=None):
    """All-gathers row blocks of unequal height ([n_local, ...] per rank, contiguous ranges in rank
    order) into one [n_total, ...] tensor on every rank.  Pads to the largest shard so that a
    single fixed-size all_gather moves the data (NCCL all_gather_into_tensor on GPU)."""
    import torch
    dist = _dist()
    rank, world = world_info()
    if world == 1:
        return local
    counts = shard_counts(n_total, world)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * mx:r * mx + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def sharded_walks(handle, p, q, walk_length, starts_all, seed, gather=True, torch_device=None):
    """One walk per entry of starts_all (same array on every rank); this rank walks its contiguous
    slice with walk ids = global positions.  Returns the [n, L] corpus (gathered) or the local block."""
    import torch
    rank, world = world_info()
    lo, hi = shard_range(len(starts_all), rank, world)
    w, _ = handle.walks(p, q, walk_length, np.ascontiguousarray(starts_all[lo:hi]), seed=seed, walk_id_base=lo)
    if not gather or world == 1:
        return w
    dev = torch_device if torch_device is not None else torch.device("cpu")
    return gather_rows(torch.from_numpy(w).to(dev), len(starts_all)).cpu().numpy()


def sharded_simrank_topk(handle, queries_all, c, step, sample, k, seed, mode=0, torch_device=None):
    """Top-k for every query (same array on every rank), queries split in contiguous ranges."""
    import torch
    rank, world = world_info()
    lo, hi = shard_range(len(queries_all), rank, world)
    ids, sc = handle.simrank_topk(np.ascontiguousarray(queries_all[lo:hi]), c, step, sample, k, mode=mode,
                                  seed=seed, query_id_base=lo)
    if world == 1:
        return ids, sc
    dev = torch_device if torch_device is not None else torch.device("cpu")
    gi = gather_rows(torch.from_numpy(ids).to(dev), len(queries_all)).cpu().numpy()
    gs = gather_rows(torch.from_numpy(sc).to(dev), len(queries_all)).cpu().numpy()
    return gi, gs
