"""Multi-GPU plumbing for the one place this path talks across ranks.

Frames are independent, so the batch is cut into contiguous shards (one per
rank, one process per GPU) and planned with no data-path collective.  The only
exchange is a sum all-reduce of the int64 statistics vector produced by
pp_stats_batch (exact integer sums, so 1/2/4/8-rank results are identical).
torch.distributed is the transport: NCCL for CUDA tensors, gloo in the CPU
tests.
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
