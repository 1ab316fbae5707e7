"""Data-parallel plumbing of the TD update: parameters, optimizer state and
target networks are replicated on every rank, the global minibatch is split into equal
contiguous row shards, and ONE all-reduce of the flat gradient arena per optimizer sub-step
turns per-shard mean-loss gradients into the global-batch gradient (every loss of the path is
a batch mean: dqn_trainer_base.py:146-155, qrdqn_trainer.py:153-155, sac_trainer.py:242,280,
td3_trainer.py:157,184).  Adam and the Polyak update then run redundantly and identically
on all ranks -- no broadcast.  The reference itself has no collective on this path
(docs/distributed.rst:12-22 only documents the intent)."""
from typing import Tuple

import torch


def shard_rows(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`; requires world | batch so that the mean of
    shard means equals the global mean."""
    if batch % world != 0:
        raise ValueError(f"global batch {batch} is not divisible by world size {world}")
    per = batch // world
    return rank * per, (rank + 1) * per


def allreduce_mean_(flat_grad: torch.Tensor, group=None) -> float:
    """In-place SUM all-reduce of the flat gradient; returns the 1/world scale the fused Adam
    applies while it reads the gradient (no extra pass over it)."""
    import torch.distributed as dist

    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)
