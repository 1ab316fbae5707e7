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
    if net.get("kind") == "dueling":
        return {"kind": "dueling", **{k: clone_net(net[k], requires_grad)
                                      for k in ("shared", "adv", "val")}}
    return {
        "W": [w.detach().clone().requires_grad_(requires_grad) for w in net["W"]],
        "b": [x.detach().clone().requires_grad_(requires_grad) for x in net["b"]],
        "act": list(net["act"]),
    }


def net_params(net: Net) -> List[torch.Tensor]:
    if net.get("kind") == "dueling":  # registration order of DuelingQNetwork's sub-modules
        return net_params(net["shared"]) + net_params(net["adv"]) + net_params(net["val"])
    out = []
    for w, b in zip(net["W"], net["b"]):
        out += [w, b]
    return out


def mlp(net: Net, x: torch.Tensor) -> torch.Tensor:
    """FullyConnectedNetwork.forward (fully_connected_network.py:157-163); for a dueling net
    DuelingQNetwork._get_values (reagent/models/dueling_q_network.py:92-103):
    q = value + (advantage - mean over the non-batch dims of advantage)."""
    if net.get("kind") == "dueling":
        shared = mlp(net["shared"], x)
        value = mlp(net["val"], shared)          # (B, N) -- N = 1 without atoms
        raw_adv = mlp(net["adv"], shared)        # (B, A*N)
        B, N = value.shape
        adv = raw_adv.view(B, -1, N)
        # mean over ALL non-batch dims (actions and atoms), dueling_q_network.py:98-101
        q = value.view(B, 1, N) + (adv - adv.mean(dim=(1, 2), keepdim=True))
        return q.reshape(B, -1)
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


def masked_softmax(x, mask, temperature):
    """reagent/core/torch_utils.py:62-73"""
    x = x / temperature
    mmx = x - ((1.0 - mask) * 1e20)
    mmx = mmx - torch.max(mmx, dim=1, keepdim=True)[0]
    e = torch.exp(mmx) * mask
    out = e / e.sum(dim=1, keepdim=True)
    out[out != out] = 0
    return out


def dqn_cpe_losses(q: Net, reward_net: Net, qcpe: Net, qcpe_t: Net, batch, *, gamma, temperature,
                   num_actions, maxq=True, loss="mse", discount_src=None):
    """_calculate_cpes (reagent/training/dqn_trainer_base.py:332-452): (reward loss, CPE
    q-value loss).  `batch["metrics"]` (B, M-1) are the extra metrics, may be absent."""
    A = num_actions
    mrc = batch["reward"]
    if batch.get("metrics") is not None and batch["metrics"].shape[1] > 0:
        mrc = torch.cat((batch["reward"], batch["metrics"]), dim=1)
    M = mrc.shape[1]
    offsets = torch.arange(0, M * A, A, dtype=torch.long)
    logged = torch.argmax(batch["action"], dim=1, keepdim=True)
    with torch.no_grad():
        next_scores = mlp(q, batch["next_state"])  # dqn_trainer.py:268 (after the q step)
    mask = (batch["possible_next_actions_mask"] if maxq else batch["next_action"]).float()
    prop = masked_softmax(next_scores, mask, temperature)
    discount = torch.full_like(batch["reward"], gamma)
    if discount_src is not None:
        discount = torch.pow(gamma, discount_src.float())
    not_done = batch["not_terminal"].float()
    r_est = mlp(reward_net, batch["state"]).gather(1, offsets + logged)
    reward_loss = F.mse_loss(r_est, mrc)
    metric_q = mlp(qcpe, batch["state"]).gather(1, offsets + logged)
    chunks = torch.chunk(mlp(qcpe_t, batch["next_state"]).detach(), M, dim=1)
    tgt = []
    for i, per_metric in enumerate(chunks):
        nq = torch.sum(per_metric * prop, 1, keepdim=True) * not_done
        tgt.append(mrc[:, i:i + 1] + discount * nq)
    tgt = torch.cat(tgt, dim=1)
    fn = F.mse_loss if loss == "mse" else F.smooth_l1_loss
    return reward_loss, fn(metric_q, tgt), prop


def dqn_cpe_update(q, reward_net, adam_r, qcpe, qcpe_t, adam_c, batch, *, tau, **kw):
    """The two CPE optimizer steps of one DQNTrainer update and the soft update of the CPE
    target; call AFTER dqn_update's Adam step of q (dqn_trainer.py:256-304).  Returns
    (reward_loss, cpe_loss, reward grads, cpe grads)."""
    pr, pc = net_params(reward_net), net_params(qcpe)
    rl, cl, _ = dqn_cpe_losses(q, reward_net, qcpe, qcpe_t, batch, **kw)
    gr = [g.detach().clone() for g in torch.autograd.grad(rl, pr)]
    adam_r.step(pr, gr)
    gc = [g.detach().clone() for g in torch.autograd.grad(cl, pc)]
    adam_c.step(pc, gc)
    soft_update(qcpe_t, qcpe, tau)
    return float(rl.detach()), float(cl.detach()), gr, gc


# ---------------------------------------------------------------------------
# Gaussian actor head (reagent/models/actor.py:169-261)
# ---------------------------------------------------------------------------
LOG_PROB_MIN, LOG_PROB_MAX = -2.0, 2.0
_ACT_EPS = 1e-6
_CONST = math.log(math.sqrt(2 * math.pi))


def gaussian_log_prob(actor: Net, state, squashed_action):
    """GaussianFullyConnectedActor.get_log_prob (actor.py:233-261)."""
    out = mlp(actor, state)
    A = out.shape[1] // 2
    loc, scale_log = out[:, :A], out[:, A:].clamp(LOG_PROB_MIN, LOG_PROB_MAX)
    raw_action = torch.atanh(squashed_action)
    r = (raw_action - loc) / scale_log.exp()
    log_prob = -(r ** 2) / 2 - scale_log - _CONST
    squash_correction = (1 - squashed_action ** 2 + _ACT_EPS).log()
    return torch.sum(log_prob - squash_correction, dim=1).reshape(-1, 1)


def gaussian_actor_forward(actor: Net, state, noise):
    """GaussianFullyConnectedActor.forward with the randn_like draw injected (actor.py:215-231)."""
    out = mlp(actor, state)
    A = out.shape[1] // 2
    loc, scale_log = out[:, :A], out[:, A:].clamp(LOG_PROB_MIN, LOG_PROB_MAX)
    raw_action = loc + noise * scale_log.exp()
    squashed = torch.clamp(torch.tanh(raw_action), -1.0 + _ACT_EPS, 1.0 - _ACT_EPS)
    return squashed, gaussian_log_prob(actor, state, squashed)


def critic(q: Net, state, action):
    """FullyConnectedCritic.forward (reagent/models/critic.py:76-92)."""
    return mlp(q, torch.cat((state, action), dim=-1))


def _grad_step(loss, net_or_params, adam):
    params = net_or_params if isinstance(net_or_params, list) else net_params(net_or_params)
    grads = torch.autograd.grad(loss, params)
    grads = [g.detach().clone() for g in grads]
    adam.step(params, grads)
    return grads


class SacState:
    """Mutable state of the restated SACTrainer (twin or single critic, no value net)."""

    def __init__(self, actor, q1, q2, *, lr=1e-3, entropy_temperature=0.01, learn_alpha=True,
                 target_entropy=-1.0):
        self.actor, self.q1, self.q2 = actor, q1, q2
        self.q1t = clone_net(q1)
        self.q2t = None if q2 is None else clone_net(q2)
        for n in (actor, q1, q2):
            if n is not None:
                for p in net_params(n):
                    p.requires_grad_(True)
        self.alpha = entropy_temperature
        self.learn_alpha = learn_alpha
        self.target_entropy = target_entropy
        self.adam_q1 = AdamState(net_params(q1), lr=lr)
        self.adam_q2 = None if q2 is None else AdamState(net_params(q2), lr=lr)
        self.adam_actor = AdamState(net_params(actor), lr=lr)
        if learn_alpha:
            # float64, as torch.tensor([np.log(x)]) is in the reference (sac_trainer.py:122-126)
            self.log_alpha = torch.tensor([math.log(entropy_temperature)], dtype=torch.float64,
                                          requires_grad=True)
            self.adam_alpha = AdamState([self.log_alpha], lr=lr)


def sac_update(st: SacState, batch, noise_next, noise_cur, *, gamma, tau,
               backprop_through_log_prob=True):
    """One SACTrainer update (reagent/training/sac_trainer.py:195-385, value_network=None).
    Returns dict(losses=[q1, q2, actor, alpha], grads={...})."""
    state, action = batch["state"], batch["action"]
    reward, not_done = batch["reward"], batch["not_terminal"].float()
    # --- target (:214-239) ---
    a_next, _ = gaussian_actor_forward(st.actor, batch["next_state"], noise_next)
    next_v = critic(st.q1t, batch["next_state"], a_next)
    if st.q2 is not None:
        next_v = torch.min(next_v, critic(st.q2t, batch["next_state"], a_next))
    log_prob_a = gaussian_log_prob(st.actor, batch["next_state"], a_next).clamp(
        LOG_PROB_MIN, LOG_PROB_MAX)
    next_v = (next_v - st.alpha * log_prob_a).float()
    discount = torch.full_like(reward, gamma)
    target = (reward + discount * next_v * not_done) if gamma > 0.0 else reward
    target = target.detach()
    out = {"losses": [], "grads": {}, "target": target}
    # --- critics (:241-248) ---
    q1_loss = F.mse_loss(critic(st.q1, state, action), target)
    out["grads"]["q1"] = _grad_step(q1_loss, st.q1, st.adam_q1)
    out["losses"].append(float(q1_loss))
    if st.q2 is not None:
        q2_loss = F.mse_loss(critic(st.q2, state, action), target)
        out["grads"]["q2"] = _grad_step(q2_loss, st.q2, st.adam_q2)
        out["losses"].append(float(q2_loss))
    # --- actor (:254-308), sees the updated critics ---
    a_cur, logp = gaussian_actor_forward(st.actor, state, noise_cur)
    min_q = critic(st.q1, state, a_cur)
    if st.q2 is not None:
        min_q = torch.min(min_q, critic(st.q2, state, a_cur))
    actor_log_prob = logp.clamp(LOG_PROB_MIN, LOG_PROB_MAX)
    if not backprop_through_log_prob:
        actor_log_prob = actor_log_prob.detach()
    actor_loss = (st.alpha * actor_log_prob - min_q).mean()
    out["grads"]["actor"] = _grad_step(actor_loss, st.actor, st.adam_actor)
    out["losses"].append(float(actor_loss))
    # --- alpha (:311-322) ---
    if st.learn_alpha:
        alpha_loss = -(
            (st.log_alpha * (logp.clamp(LOG_PROB_MIN, LOG_PROB_MAX) + st.target_entropy).detach())
            .mean())
        out["grads"]["alpha"] = _grad_step(alpha_loss, [st.log_alpha], st.adam_alpha)
        out["losses"].append(float(alpha_loss))
        st.alpha = st.log_alpha.detach().exp()
    # --- soft update (:383-385) ---
    soft_update(st.q1t, st.q1, tau)
    if st.q2 is not None:
        soft_update(st.q2t, st.q2, tau)
    return out


class Td3State:
    def __init__(self, actor, q1, q2, *, lr=1e-3):
        self.actor, self.q1, self.q2 = actor, q1, q2
        self.actor_t, self.q1t = clone_net(actor), clone_net(q1)
        self.q2t = None if q2 is None else clone_net(q2)
        for n in (actor, q1, q2):
            if n is not None:
                for p in net_params(n):
                    p.requires_grad_(True)
        self.adam_q1 = AdamState(net_params(q1), lr=lr)
        self.adam_q2 = None if q2 is None else AdamState(net_params(q2), lr=lr)
        self.adam_actor = AdamState(net_params(actor), lr=lr)


def td3_update(st: Td3State, batch, noise_next, batch_idx, *, gamma, tau, noise_variance=0.2,
               noise_clip=0.5, delayed_policy_update=2):
    """One TD3Trainer update (reagent/training/td3_trainer.py:125-199)."""
    state, action = batch["state"], batch["action"]
    with torch.no_grad():  # :138-153
        next_actor = mlp(st.actor_t, batch["next_state"])
        noise = noise_next * noise_variance
        next_actor = (next_actor + noise.clamp(-noise_clip, noise_clip)).clamp(-1.0, 1.0)
        next_q = critic(st.q1t, batch["next_state"], next_actor)
        if st.q2 is not None:
            next_q = torch.min(next_q, critic(st.q2t, batch["next_state"], next_actor))
        target = batch["reward"] + gamma * next_q * batch["not_terminal"].float()
    out = {"losses": [], "grads": {}, "target": target}
    q1_loss = F.mse_loss(critic(st.q1, state, action), target)
    out["grads"]["q1"] = _grad_step(q1_loss, st.q1, st.adam_q1)
    out["losses"].append(float(q1_loss))
    if st.q2 is not None:
        q2_loss = F.mse_loss(critic(st.q2, state, action), target)
        out["grads"]["q2"] = _grad_step(q2_loss, st.q2, st.adam_q2)
        out["losses"].append(float(q2_loss))
    if batch_idx % delayed_policy_update == 0:  # :181-194
        actor_loss = -(critic(st.q1, state, mlp(st.actor, state)).mean())
        out["grads"]["actor"] = _grad_step(actor_loss, st.actor, st.adam_actor)
        out["losses"].append(float(actor_loss))
        soft_update(st.q1t, st.q1, tau)
        if st.q2 is not None:
            soft_update(st.q2t, st.q2, tau)
        soft_update(st.actor_t, st.actor, tau)
    else:
        out["losses"].append(None)
    return out


# ---------------------------------------------------------------------------
# QR-DQN (reagent/training/qrdqn_trainer.py:108-194, :210-218)
# ---------------------------------------------------------------------------
def qrdqn_loss(q: Net, qt: Net, batch, *, gamma, num_atoms, double_q=True, maxq=True,
               discount_src=None, reward_boost=None):
    reward, action = batch["reward"], batch["action"]
    B, A, N = reward.shape[0], action.shape[1], num_atoms
    if reward_boost is not None:
        reward = reward + torch.sum(action.float() * reward_boost, dim=1, keepdim=True)
    discount = torch.full_like(reward, gamma)
    if discount_src is not None:
        discount = torch.pow(gamma, discount_src.float())
    not_done = batch["not_terminal"].float()
    quantiles = ((0.5 + torch.arange(N).float()) / float(N)).view(1, -1)  # :70-73
    next_qf = mlp(qt, batch["next_state"]).view(B, A, N)  # :125
    if maxq:
        next_q_values = (mlp(q, batch["next_state"]).view(B, A, N) if double_q else next_qf).mean(dim=2)
        qv = next_q_values + ACTION_NOT_POSSIBLE_VAL * (1 - batch["possible_next_actions_mask"].float())
        next_action = qv.argmax(1)  # :210-214
        next_qf = next_qf[range(B), next_action.reshape(-1)]
    else:
        next_action = None
        next_qf = (next_qf * batch["next_action"].unsqueeze(-1)).sum(1)
    target_Q = (reward + discount * not_done * next_qf).detach()  # :142
    current_qf = mlp(q, batch["state"]).view(B, A, N)
    all_q = current_qf.mean(2).detach()
    current_qf = (current_qf * action.unsqueeze(-1)).sum(1)  # :149
    td = target_Q.t().unsqueeze(-1) - current_qf  # (N, B, N), :152
    huber = torch.where(td.abs() < 1, 0.5 * td.pow(2), td.abs() - 0.5)  # :217-218
    loss = (huber * (quantiles - (td.detach() < 0).float()).abs()).mean()  # :153-155
    return loss, {"next_action": next_action, "all_q": all_q, "target": target_Q}


def qrdqn_update(q: Net, qt: Net, adam: AdamState, batch, *, gamma, tau, num_atoms, **kw):
    params = net_params(q)
    loss, aux = qrdqn_loss(q, qt, batch, gamma=gamma, num_atoms=num_atoms, **kw)
    grads = [g.detach().clone() for g in torch.autograd.grad(loss, params)]
    adam.step(params, grads)
    soft_update(qt, q, tau)
    return float(loss.detach()), grads, aux


# ---------------------------------------------------------------------------
# ParametricDQN (reagent/training/parametric_dqn_trainer.py:109-214)
# ---------------------------------------------------------------------------
def pdqn_update(q: Net, qt: Net, adam: AdamState, batch, *, gamma, tau, double_q=True, maxq=True,
                loss="mse", discount_src=None, reward_net: Optional[Net] = None,
                adam_r: Optional[AdamState] = None):
    """One ParametricDQNTrainer update.  batch: state/next_state (B,S), action/next_action (B,Ad),
    possible_next_actions (B*M,Ad), possible_next_actions_mask (B,M), reward/not_terminal (B,1).
    Returns (td_loss, reward_loss | None, grads of q)."""
    reward, not_terminal = batch["reward"], batch["not_terminal"].float()
    B = reward.shape[0]
    discount = torch.full_like(reward, gamma)
    if discount_src is not None:
        discount = torch.pow(gamma, discount_src.float())
    with torch.no_grad():
        if maxq:
            pna = batch["possible_next_actions"]
            M = pna.shape[0] // B
            tiled = batch["next_state"].repeat_interleave(M, dim=0)  # get_tiled_batch
            all_q, all_qt = critic(q, tiled, pna), critic(qt, tiled, pna)
            mask = batch["possible_next_actions_mask"].float()
            qv = all_q.reshape(mask.shape) + ACTION_NOT_POSSIBLE_VAL * (1 - mask)
            qtv = all_qt.reshape(mask.shape) + ACTION_NOT_POSSIBLE_VAL * (1 - mask)
            if double_q:
                _, idx = torch.max(qv, dim=1, keepdim=True)
                next_q = torch.gather(qtv, 1, idx)
            else:
                next_q, _ = torch.max(qtv, dim=1, keepdim=True)
        else:
            next_q = critic(qt, batch["next_state"], batch["next_action"])
    target = reward + not_terminal * discount * next_q
    fn = F.mse_loss if loss == "mse" else F.smooth_l1_loss
    params = net_params(q)
    td = fn(critic(q, batch["state"], batch["action"]), target)
    grads = [g.detach().clone() for g in torch.autograd.grad(td, params)]
    adam.step(params, grads)
    rl = None
    if reward_net is not None:
        mrc = reward if batch.get("metrics") is None else torch.cat((reward, batch["metrics"]), dim=1)
        est = critic(reward_net, batch["state"], batch["action"])
        rloss = F.mse_loss(est.squeeze(-1), mrc.squeeze(-1))
        pr = net_params(reward_net)
        adam_r.step(pr, [g.detach().clone() for g in torch.autograd.grad(rloss, pr)])
        rl = float(rloss.detach())
    soft_update(qt, q, tau)
    return float(td.detach()), rl, grads


# ---------------------------------------------------------------------------
# C51 (reagent/training/c51_trainer.py:98-173, reagent/models/categorical_dqn.py:28-35)
# ---------------------------------------------------------------------------
def c51_loss(q: Net, qt: Net, batch, *, gamma, num_atoms, qmin, qmax, double_q=True, maxq=True,
             discount_src=None, reward_boost=None):
    reward, action = batch["reward"], batch["action"]
    B, A, N = reward.shape[0], action.shape[1], num_atoms
    support = torch.linspace(qmin, qmax, N)
    scale_support = (qmax - qmin) / (N - 1.0)
    if reward_boost is not None:
        reward = reward + torch.sum(action.float() * reward_boost, dim=1, keepdim=True)
    discount = torch.full_like(reward, gamma)
    if discount_src is not None:
        discount = torch.pow(gamma, discount_src.float())
    not_terminal = batch["not_terminal"].float()
    log_dist = lambda net, x: F.log_softmax(mlp(net, x).view(B, A, N), -1)  # noqa: E731
    with torch.no_grad():
        next_dist = log_dist(qt, batch["next_state"]).exp()
        if maxq:
            if double_q:
                next_q = (log_dist(q, batch["next_state"]).exp() * support).sum(2)
            else:
                next_q = (next_dist * support).sum(2)
            mask = batch["possible_next_actions_mask"].float()
            next_action = (next_q + ACTION_NOT_POSSIBLE_VAL * (1 - mask)).argmax(1)
            next_dist = next_dist[range(B), next_action.reshape(-1)]
        else:
            next_dist = (next_dist * batch["next_action"].unsqueeze(-1)).sum(1)
        target_Q = (reward + discount * not_terminal * support).clamp(qmin, qmax)
        b = (target_Q - qmin) / scale_support
        lo, up = b.floor().to(torch.int64), b.ceil().to(torch.int64)
        lo[(up > 0) * (lo == up)] -= 1
        up[(lo < (N - 1)) * (lo == up)] += 1
        m = torch.zeros_like(next_dist)
        m.scatter_add_(dim=1, index=lo, src=next_dist * (up.float() - b))
        m.scatter_add_(dim=1, index=up, src=next_dist * (b - lo.float()))
    ld = (log_dist(q, batch["state"]) * action.unsqueeze(-1)).sum(1)
    return -(m * ld).sum(1).mean()


def c51_update(q: Net, qt: Net, adam: AdamState, batch, *, gamma, tau, **kw):
    params = net_params(q)
    loss = c51_loss(q, qt, batch, gamma=gamma, **kw)
    grads = [g.detach().clone() for g in torch.autograd.grad(loss, params)]
    adam.step(params, grads)
    soft_update(qt, q, tau)
    return float(loss.detach()), grads
