"""Replicate sharding for one-process-per-GPU runs.

Permutation replicates are independent and the only cross-replicate state is the additive exceedance counter, so rank r
of W takes a contiguous replicate range, runs it with `perm_offset` = its first global replicate index (the Philox
counter is the global index, so the union of the shards reproduces the single-GPU labels) and the int64 counters are
merged with ONE all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests of this host logic).
"""
from __future__ import annotations


def shard_range(iterations: int, rank: int, world: int) -> tuple[int, int]:
    """(first replicate, number of replicates) of `rank`; same split as the in-library multi-GPU path (ceil blocks)."""
    per = -(-int(iterations) // int(world))
    begin = min(int(iterations), per * rank)
    end = min(int(iterations), per * (rank + 1))
    return begin, end - begin


def merge_counts(counts, group=None):
    """In-place sum of an int64 counter tensor over the process group (the path's single exchange step)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def run_sharded(engine, iterations, counts_dev, *, seed=42, score_type="geometric_mean", batch_size=0, stream=None,
                group=None, precision="fp64", rel_tol=0.0, band_dev=None):
    """Run this rank's shard of `iterations` replicates on `engine` (its GPU), leave the merged int64 counters in the
    torch CUDA tensor `counts_dev` on every rank (and the band counters of precision="fp32_fast" in `band_dev`, if given).
    Returns the engine stats of the local shard."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    begin, n = shard_range(iterations, rank, world)
    stats = None
    if n > 0:
        stats = engine.run(n, seed=seed, score_type=score_type, perm_offset=begin, batch_size=batch_size,
                           counts_device_ptr=counts_dev.data_ptr(), want_host_counts=False, stream=stream,
                           precision=precision, rel_tol=rel_tol,
                           band_device_ptr=None if band_dev is None else band_dev.data_ptr())["stats"]
    else:
        counts_dev.zero_()
        if band_dev is not None:
            band_dev.zero_()
    merge_counts(counts_dev, group)
    if band_dev is not None:
        merge_counts(band_dev, group)
    return stats
