"""Host-side plumbing of data-parallel training (new functionality: the reference is single-device,
train.py:36-43).  One process per GPU; `torch.distributed` is only the bootstrap channel -- the gradient
all-reduce itself is issued by the native step on its own NCCL communicator (csrc/trainer.cu)."""
from typing import Callable, Tuple

import numpy as np


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r trains on pairs [r*B/n, (r+1)*B/n) of the global batch; equal shards only, because
    every rank's loss is a mean over ITS pairs and the gradients are averaged with equal weights."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_batch(a: np.ndarray, b: np.ndarray, rank: int, world: int):
    lo, hi = shard_bounds(len(a), rank, world)
    return a[lo:hi], b[lo:hi]


def exchange_unique_id(make_id: Callable[[], bytes], dist, device="cpu") -> bytes:
    """Rank 0 creates the 128-byte NCCL unique id (cg_comm_unique_id); everyone receives it over the bootstrap
    process group (any backend)."""
    import torch
    rank = dist.get_rank()
    if rank == 0:
        raw = make_id()
        if len(raw) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def mean_over_ranks(values, dist, device="cpu"):
    """All-reduce-mean of a few host scalars (the 6 metrics when a caller wants global numbers)."""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t)
    return (t / dist.get_world_size()).tolist()
