"""K2 time as a function of network shape (production build): T = a + b * steps + c * chunks.
Separates the per-layer hand-over cost from the per-chunk streaming cost without instrumenting
the kernel.  Usage: python profiles/k2_shapes.py"""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reagent_b200 import _lib
from reagent_b200.core import types as rlt
from reagent_b200.core.parameters import EvaluationParameters, RLParameters
from reagent_b200.models import FullyConnectedDQN
from reagent_b200.optimizer import Optimizer__Union
from reagent_b200.training import DQNTrainer

dev = torch.device("cuda", 0)
B, A = 4096, 16


def timeit(fn, n=300):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def count(S, hidden):
    dims = [S] + hidden + [A]
    L = len(dims) - 1
    cd = lambda a, b: -(-a // b)
    fwd = sum(cd(dims[l + 1], 128) * cd(dims[l], 32) for l in range(L))
    bwd = sum(cd(dims[l], 128) * cd(dims[l + 1], 32) for l in range(1, L))
    return 3 * L + (L - 1), 3 * fwd + bwd


rows = []
for S, hidden in ((128, [256, 128]), (128, [128, 128]), (128, [128]), (128, [256]), (64, [64, 64]), (128, [256, 256, 128]),
                  (256, [256, 128]), (128, [128, 128, 128, 128])):
    torch.manual_seed(0)
    q = FullyConnectedDQN(S, A, hidden, ["relu"] * len(hidden)); qt = q.get_target_network()
    t = DQNTrainer(q.to(dev), qt.to(dev), actions=[str(i) for i in range(A)],
                   rl=RLParameters(gamma=0.99, target_update_rate=0.005, q_network_loss="huber"),
                   optimizer=Optimizer__Union.default(lr=1e-3),
                   evaluation=EvaluationParameters(calc_cpe_in_training=False)).to(dev)
    act = torch.randint(A, (B,), device=dev)
    batch = rlt.DiscreteDqnInput(
        state=rlt.FeatureData(torch.randn(B, S, device=dev)), next_state=rlt.FeatureData(torch.randn(B, S, device=dev)),
        reward=torch.randn(B, 1, device=dev), time_diff=None, step=None,
        not_terminal=torch.ones(B, 1, device=dev), action=torch.nn.functional.one_hot(act, A).float(),
        next_action=torch.nn.functional.one_hot(act, A).float(),
        possible_actions_mask=torch.ones(B, A, device=dev), possible_next_actions_mask=torch.ones(B, A, device=dev),
        extras=rlt.ExtraData())
    t._td_step(batch)
    qd, qtd, a, wsc, keep, pack = t._last_td_call
    assert pack is not None
    st = _lib.cur_stream()
    us = timeit(lambda: _lib.lib().rb200_dqn_td_step_tc(qd, qtd, a, wsc, pack.data_ptr(), pack.numel(), 1, st))
    steps, chunks = count(S, hidden)
    rows.append((steps, chunks, us))
    print(json.dumps({"S": S, "hidden": hidden, "steps": steps, "chunks32": chunks, "k2_us": round(us, 2)}), flush=True)
M = np.array([[1.0, r[0], r[1]] for r in rows]); y = np.array([r[2] for r in rows])
coef, res, *_ = np.linalg.lstsq(M, y, rcond=None)
print(json.dumps({"fit_us": {"fixed": round(coef[0], 2), "per_step": round(coef[1], 3), "per_chunk32": round(coef[2], 3)},
                  "max_abs_residual_us": round(float(np.abs(M @ coef - y).max()), 2)}))
