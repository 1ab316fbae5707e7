"""world_size-2 gloo test of the data-parallel host logic: shard means all-reduced through
reagent_b200.training.data_parallel equal the full-batch gradient, and the replicated Adam
step keeps the ranks bit-identical (CPU oracle arithmetic, no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import td_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    from reagent_b200.training.data_parallel import allreduce_mean_, shard_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    gen = torch.Generator().manual_seed(0)
    S, A, B = 12, 5, 64
    q = O.make_net([S, 24, 20, A], ["relu", "relu", "linear"], gen)
    qt = O.clone_net(q)
    act = torch.randint(A, (B,), generator=gen)
    batch = dict(state=torch.randn(B, S, generator=gen), next_state=torch.randn(B, S, generator=gen),
                 reward=torch.randn(B, 1, generator=gen),
                 not_terminal=(torch.rand(B, 1, generator=gen) > 0.2).float(),
                 action=torch.nn.functional.one_hot(act, A).float(),
                 possible_next_actions_mask=torch.ones(B, A), next_action=None)
    # full-batch gradient (what a single GPU computes)
    qf = O.clone_net(q, requires_grad=True)
    loss, _ = O.dqn_td_loss(qf, qt, batch, gamma=0.9, loss="huber")
    full = torch.cat([g.reshape(-1) for g in torch.autograd.grad(loss, O.net_params(qf))])
    # this rank's shard
    lo, hi = shard_rows(B, rank, world)
    shard = {k: (v[lo:hi] if v is not None else None) for k, v in batch.items()}
    qs = O.clone_net(q, requires_grad=True)
    ls, _ = O.dqn_td_loss(qs, qt, shard, gamma=0.9, loss="huber")
    flat = torch.cat([g.reshape(-1) for g in torch.autograd.grad(ls, O.net_params(qs))])
    scale = allreduce_mean_(flat)
    flat *= scale
    err = float((flat - full).abs().max() / full.abs().max())
    # replicated Adam step on the averaged gradient -> identical parameters on every rank
    adam = O.AdamState(O.net_params(qs), lr=1e-2)
    grads, off = [], 0
    for p in O.net_params(qs):
        grads.append(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    adam.step(O.net_params(qs), grads)
    w = torch.cat([p.detach().reshape(-1) for p in O.net_params(qs)])
    gathered = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        ret["err"] = err
        ret["same"] = same
    dist.destroy_process_group()


def test_shard_rows_validation():
    from reagent_b200.training.data_parallel import shard_rows

    assert shard_rows(4096, 3, 8) == (1536, 2048)
    with pytest.raises(ValueError):
        shard_rows(10, 0, 4)


def test_two_rank_gradient_average_matches_full_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["err"] < 1e-6
    assert ret["same"]
