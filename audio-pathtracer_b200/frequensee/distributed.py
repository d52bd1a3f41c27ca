"""Multi-GPU plumbing: one process per GPU, the global work range g = source * n_paths + i is cut
into contiguous shards, every rank traces its shard against a replicated BVH into a private
u64 histogram, and the per-rank histograms are combined with ONE integer reduce (NCCL over
NVLink on GPUs, gloo in the CPU tests).  Integer addition is associative, so the result is
bit-identical for every rank count and every reduction topology.

The reference has no counterpart (single process, SUB.cpp:215-230 is a serial loop); the
independent unit is one iteration of that loop.
"""
import numpy as np


def shard_range(total, rank, world):
    """contiguous [first, first+count) of `total` work items for `rank` of `world`"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(total), int(world))
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def reduce_histogram(hist_tensor, dst=0, all_ranks=False):
    """Sum an int64/uint64-bit-pattern histogram tensor over the default process group.
    int64 two's-complement addition is the same bit operation as uint64 addition."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return hist_tensor
    if all_ranks:
        dist.all_reduce(hist_tensor, op=dist.ReduceOp.SUM)
    else:
        dist.reduce(hist_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return hist_tensor


def trace_sharded(trace_range_fn, n_sources, n_paths, rank, world):
    """Runs `trace_range_fn(g_first, g_count)` for this rank's shard; returns what it returns."""
    first, count = shard_range(n_sources * n_paths, rank, world)
    return trace_range_fn(first, count)
