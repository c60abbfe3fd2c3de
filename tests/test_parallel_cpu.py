"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic, and of the property the native DP step relies
on: with equal shards, mean-of-per-rank gradients == the global-batch gradient (losses are per-batch means and
instance norm is per sample)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from cyclegan_cat_b200.parallel import exchange_unique_id, mean_over_ranks, shard_batch
    from oracle.train import OracleCycleGan, synthetic_batch
    from tests import common as C
    try:
        # 1. the NCCL id travels from rank 0 to everyone
        raw = exchange_unique_id(lambda: bytes(range(128)), dist)
        assert raw == bytes(range(128))
        # 2. equal shards, gradient mean == global-batch gradient (oracle, fp64)
        a, b = synthetic_batch(4, 32)
        o = OracleCycleGan(C.SMALL_RESNET, C.SMALL_SIMPLE, dtype=torch.float64)
        sa, sb = shard_batch(a, b, rank, world)
        assert len(sa) == 2
        m_local, g_local, _ = o.gradients(sa, sb)
        flat = torch.cat([g.reshape(-1) for net in ("g_AB", "g_BA", "d_A", "d_B") for g in g_local[net]])
        dist.all_reduce(flat)
        flat /= world
        m_global, g_global, _ = o.gradients(a, b)
        ref = torch.cat([g.reshape(-1) for net in ("g_AB", "g_BA", "d_A", "d_B") for g in g_global[net]])
        err = float((flat - ref).norm() / ref.norm())
        means = mean_over_ranks([m_local[k] for k in ("gAB_loss", "dA_loss")], dist)
        derr = abs(means[0] - m_global["gAB_loss"]) + abs(means[1] - m_global["dA_loss"])
        if rank == 0:
            out.put((err, derr))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_data_parallel_equivalence():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    err, derr = out.get(timeout=10)
    assert err < 1e-10, err
    assert derr < 1e-10, derr


def test_shard_bounds():
    from cyclegan_cat_b200.parallel import shard_bounds
    assert [shard_bounds(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]      # C4: global 64 on 8 GPUs
    assert shard_bounds(64, 1, 2) == (32, 64)
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)
