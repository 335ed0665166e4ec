"""Multi-GPU plumbing for the one place this path talks across ranks.

Frames are independent, so the batch is cut into contiguous shards (one per
rank, one process per GPU) and planned with no data-path collective.  The only
exchange is the final reduction of the statistics: a sum of the int64 vector of
pp_stats_batch (exact integer sums, so 1/2/4/8-rank results are identical) and
a min / max of the f64 vector of pp_fstats_batch.  On GPUs that reduction is
pp_stats_reduce in the C library (NCCL); the functions below are the same rules
over torch.distributed (gloo in the CPU tests) and on single vectors (the
N-rank == 1-rank check of bench.py).
"""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`: [r*N/G, (r+1)*N/G)."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


def allreduce_stats(stats):
    """In-place SUM all-reduce of an int64 statistics tensor over the default
    process group (no-op when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    assert stats.dtype == torch.int64
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def merge_fstats(a, b, n_min: int):
    """The reduction rule of the f64 statistics (pp.h, PP_FSTAT_*): element-wise minimum over
    the first n_min entries, maximum over the rest.  Works on torch tensors and numpy arrays."""
    import numpy as np
    if isinstance(a, np.ndarray):
        out = a.copy()
        out[:n_min] = np.minimum(a[:n_min], b[:n_min])
        out[n_min:] = np.maximum(a[n_min:], b[n_min:])
        return out
    import torch
    out = a.clone()
    out[:n_min] = torch.minimum(a[:n_min], b[:n_min])
    out[n_min:] = torch.maximum(a[n_min:], b[n_min:])
    return out


def allreduce_fstats(fstats, n_min: int):
    """In-place MIN / MAX all-reduce of an f64 statistics tensor over the default process group
    (what pp_stats_reduce does with ncclMin / ncclMax)."""
    import torch
    import torch.distributed as dist
    assert fstats.dtype == torch.float64
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        lo, hi = fstats[:n_min].clone(), fstats[n_min:].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        fstats[:n_min], fstats[n_min:] = lo, hi
    return fstats
