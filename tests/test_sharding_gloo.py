"""CPU test of the N > 1 host logic with world_size = 2 over gloo: contiguous replicate shards + one all-reduce of the
int64 counters reproduce the single-process counts. The per-shard counts come from the CPU oracle here (no GPU in this
container); on the GPU box the same logic runs with the engine and NCCL (bench.py, tests/test_parity_gpu.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scdiffcom_b200.sharding import merge_counts, shard_range
from tests import helpers


def test_shard_range_partitions_the_replicates():
    for iters in (1, 7, 128, 1000, 100_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(iters, r, world) for r in range(world)]
            assert sum(n for _, n in spans) == iters
            pos = 0
            for b, n in spans:
                assert n == 0 or b == pos
                pos += n


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_c
    case = helpers.liver_case(helpers.load_liver(), "cond")
    prob = case["prob"]
    begin, n = shard_range(total, rank, world)
    cond, ct = oracle_c.philox_labels(prob, n, seed=9, perm_offset=begin)
    counts, _ = oracle_c.perm_counts(prob, ct, cond, nthreads=2)
    t = torch.from_numpy(np.nan_to_num(counts).astype(np.int64))
    merge_counts(t)
    if rank == 0:
        np.save(os.path.join(out_dir, "merged.npy"), t.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_merge_equals_single_process(tmp_path):
    from oracle import oracle_c
    total, world, port = 90, 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    merged = np.load(tmp_path / "merged.npy")
    prob = helpers.liver_case(helpers.load_liver(), "cond")["prob"]
    cond, ct = oracle_c.philox_labels(prob, total, seed=9)
    whole, _ = oracle_c.perm_counts(prob, ct, cond, nthreads=2)
    assert np.array_equal(merged, np.nan_to_num(whole).astype(np.int64))
