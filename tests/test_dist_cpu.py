"""Multi-rank host logic on CPU (gloo, world_size 2): sharding covers the batch exactly once, counter-based inputs
make a shard independent of the others, and the max-reduction gives every rank the single-process answer bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from exahype_b200.dist import PatchSharding, TimestepReducer, admissible_dt  # noqa: E402


def test_sharding_partitions_the_batch():
    for B in (0, 1, 7, 1000, 32768):
        for R in (1, 2, 3, 8):
            shards = [PatchSharding(B, R, r) for r in range(R)]
            assert shards[0].first == 0 and shards[-1].last == B
            assert all(a.last == b.first for a, b in zip(shards, shards[1:]))
            assert sum(s.count for s in shards) == B
            assert max(s.count for s in shards) - min(s.count for s in shards) <= 1
    with pytest.raises(ValueError):
        PatchSharding(10, 2, 2)
    assert admissible_dt(2.0, 0.1, 0.5) == 0.025 and admissible_dt(0.0, 0.1) == float("inf")


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    cfg = O.OracleConfig(dim=3, patch_size=8, halo=1, n_real=5, n_aux=0)
    B = 21
    shard = PatchSharding(B, world, rank)
    q = O.fill_synthetic(cfg, shard.count, first_patch=shard.first)     # no communication needed for the input
    lam, lmax = O.step(cfg, q, 0.01)                                     # stand-in for the per-GPU kernel
    t = torch.tensor([lmax], dtype=torch.float64)
    red = TimestepReducer(world, rank, use_nccl=False)
    red.allreduce_max(t)
    ret[rank] = (float(t.item()), O.fnv1a64(q), shard.first, shard.last)
    dist.destroy_process_group()


def test_two_ranks_equal_single_process_bitwise(oracle):
    world, port = 2, 29500 + os.getpid() % 2000
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        got = dict(ret)
    cfg = oracle.OracleConfig(dim=3, patch_size=8, halo=1, n_real=5, n_aux=0)
    q = oracle.fill_synthetic(cfg, 21)
    lam, lmax = oracle.step(cfg, q, 0.01)
    assert got[0][0] == got[1][0] == float(lmax)           # every rank holds the global maximum, exactly
    for r in range(world):
        _, h, lo, hi = got[r]
        assert h == oracle.fnv1a64(q[lo:hi])               # the shard's result equals the slice of the whole batch
