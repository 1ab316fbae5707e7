"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32 + autograd) of the reference's
per-minibatch TD updates.  Never imported by the product path (reagent_b200/); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

PINNED: tests/test_oracle_golden.py checks every function here against golden vectors in
tests/golden/*.npz that oracle/make_golden.py produced by running the UNMODIFIED reference
classes (DQNTrainer, QRDQNTrainer, SACTrainer, TD3Trainer, torch.optim.Adam, SoftUpdate)
from /root/reference through oracle/ref_harness.py.

Each function cites the reference file:line it restates.  Networks are plain lists of
(W [out,in], b [out]) tensors + activation names (FullyConnectedNetwork,
reagent/models/fully_connected_network.py:101-163).
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Net = Dict  # {"W": [Tensor], "b": [Tensor], "act": [str]}

_ACT = {
    "linear": lambda x: x,
    "relu": torch.relu,
    "tanh": torch.tanh,
    "leaky_relu": lambda x: F.leaky_relu(x, 0.01),
    "sigmoid": torch.sigmoid,
    "softplus": F.softplus,
}


def make_net(dims: List[int], acts: List[str], gen: torch.Generator) -> Net:
    """Weights ~ N(0, gain/sqrt(d_in)), bias 0 (fully_connected_network.py:21-23,:122-126)."""
    W, b = [], []
    for i, a in enumerate(acts):
        try:
            gain = torch.nn.init.calculate_gain(a)
        except ValueError:
            gain = 1.0
        W.append(torch.randn(dims[i + 1], dims[i], generator=gen) * (gain * math.sqrt(1.0 / dims[i])))
        b.append(torch.zeros(dims[i + 1]))
    return {"W": W, "b": b, "act": list(acts)}


def clone_net(net: Net, requires_grad: bool = False) -> Net:
    return {
        "W": [w.detach().clone().requires_grad_(requires_grad) for w in net["W"]],
        "b": [x.detach().clone().requires_grad_(requires_grad) for x in net["b"]],
        "act": list(net["act"]),
    }


def net_params(net: Net) -> List[torch.Tensor]:
    out = []
    for w, b in zip(net["W"], net["b"]):
        out += [w, b]
    return out


def mlp(net: Net, x: torch.Tensor) -> torch.Tensor:
    """FullyConnectedNetwork.forward (fully_connected_network.py:157-163)."""
    for w, b, a in zip(net["W"], net["b"], net["act"]):
        x = _ACT[a](F.linear(x, w, b))
    return x


# ---------------------------------------------------------------------------
# optimizer steps
# ---------------------------------------------------------------------------
class AdamState:
    """torch.optim.Adam single-tensor math (what Optimizer__Union.default() builds:
    reagent/optimizer/uninferrable_optimizers.py:23-33, optimizer.py:64-85)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    @torch.no_grad()
    def step(self, params, grads):
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.t
        bc2 = 1 - b2 ** self.t
        step_size = self.lr / bc1
        bc2_sqrt = bc2 ** 0.5
        for p, g, m, v in zip(params, grads, self.m, self.v):
            if self.wd != 0:
                g = g.add(p, alpha=self.wd)
            m.lerp_(g, 1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / bc2_sqrt).add_(self.eps)
            p.addcdiv_(m, denom, value=-step_size)


@torch.no_grad()
def soft_update(target: Net, source: Net, tau: float):
    """SoftUpdate.step (reagent/optimizer/soft_update.py:47-71)."""
    for t, s in zip(net_params(target), net_params(source)):
        t.copy_(tau * s + (1.0 - tau) * t)


# ---------------------------------------------------------------------------
# DQN (reagent/training/dqn_trainer.py:166-239, dqn_trainer_base.py:33-77,216-241)
# ---------------------------------------------------------------------------
ACTION_NOT_POSSIBLE_VAL = -1e9


def dqn_td_loss(q: Net, qt: Net, batch: Dict[str, torch.Tensor], *, gamma: float,
                double_q: bool = True, maxq: bool = True, loss: str = "mse",
                discount_src: Optional[torch.Tensor] = None,
                reward_boost: Optional[torch.Tensor] = None):
    """Returns (td_loss, aux) where aux holds target / q_selected / argmax / all q(s)."""
    reward = batch["reward"]
    action = batch["action"]
    if reward_boost is not None:  # dqn_trainer_base.py:216-241
        reward = reward + torch.sum(action.float() * reward_boost, dim=1, keepdim=True)
    discount = torch.full_like(reward, gamma)  # dqn_trainer.py:166-177
    if discount_src is not None:
        discount = torch.pow(gamma, discount_src.float())
    not_done = batch["not_terminal"].float()
    with torch.no_grad():  # dqn_trainer.py:157-164
        q_next = mlp(q, batch["next_state"])
        q_next_t = mlp(qt, batch["next_state"])
    mask = (batch["possible_next_actions_mask"] if maxq else batch["next_action"]).float()
    pen = ACTION_NOT_POSSIBLE_VAL * (1 - mask)  # dqn_trainer_base.py:59-62
    qn, qnt = q_next + pen, q_next_t + pen
    if double_q:
        _, idx = torch.max(qn, dim=1, keepdim=True)
        next_q = torch.gather(qnt, 1, idx)
    else:
        next_q, idx = torch.max(qnt, dim=1, keepdim=True)
    target = reward + discount * (next_q * not_done)  # dqn_trainer.py:229-231
    all_q = mlp(q, batch["state"])
    q_sel = torch.sum(all_q * action, 1, keepdim=True)
    fn = F.mse_loss if loss == "mse" else F.smooth_l1_loss  # dqn_trainer_base.py:146-155
    td = fn(q_sel, target.detach())
    return td, {"target": target.detach(), "q_selected": q_sel.detach(), "next_idx": idx,
                "all_q": all_q.detach()}


def dqn_update(q: Net, qt: Net, adam: AdamState, batch, *, gamma, tau, **kw):
    """One full DQNTrainer update with CPE off: Adam(q) then SoftUpdate
    (dqn_trainer.py:241-304 driven by the loop in ref_harness.run_update).
    q's tensors must have requires_grad=True.  Returns (loss, grads, aux)."""
    params = net_params(q)
    for p in params:
        p.grad = None
    loss, aux = dqn_td_loss(q, qt, batch, gamma=gamma, **kw)
    loss.backward()
    grads = [p.grad.detach().clone() for p in params]
    adam.step(params, grads)
    soft_update(qt, q, tau)
    return float(loss.detach()), grads, aux
