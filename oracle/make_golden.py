"""Generate golden vectors in tests/golden/ by running the UNMODIFIED reference from
/root/reference (build container only; the files are committed because the reference does
not travel to the GPU box).

    python oracle/make_golden.py            # regenerate everything

Every .npz holds the inputs (initial weights, batch, injected noise, hyper-parameters as
0-d arrays) and the reference's outputs (losses per update, gradients of the first update,
parameters / target parameters after N updates).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.ref_harness import ref, run_update  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
N_UPDATES = 3


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, arrays, meta):
    os.makedirs(GOLDEN, exist_ok=True)
    arrays = dict(arrays)
    arrays["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


def _fc_params(module):
    """[(W, b)] of a reference network's FullyConnectedNetwork."""
    out = []
    for seq in module.fc.dnn:
        lin = seq[0]
        out.append((lin.weight, lin.bias))
    return out


def _dump_net(arrays, prefix, module):
    for i, (w, b) in enumerate(_fc_params(module)):
        arrays[f"{prefix}.W{i}"] = _np(w).copy()
        arrays[f"{prefix}.b{i}"] = _np(b).copy()


# ---------------------------------------------------------------------------
def dqn_case(name, *, B=48, S=12, A=5, sizes=(24, 20), acts=("relu", "relu"), loss="huber",
             double_q=True, maxq=True, multi_steps=None, time_diff=False, boost=None,
             random_masks=False, gamma=0.97, tau=0.05, lr=1e-2, seed=0):
    rlt = ref("reagent.core.types")
    params = ref("reagent.core.parameters")
    dqn_mod = ref("reagent.models.dqn")
    tr = ref("reagent.training.dqn_trainer")
    union = ref("reagent.optimizer.union")
    torch.manual_seed(seed)
    q = dqn_mod.FullyConnectedDQN(S, A, list(sizes), list(acts))
    # biases are 0 at init in the reference; perturb so that bias paths are exercised
    with torch.no_grad():
        for _, b in _fc_params(q):
            b.normal_(0, 0.1)
    qt = q.get_target_network()
    with torch.no_grad():
        for w, b in _fc_params(qt):
            w.add_(torch.randn_like(w) * 0.05)
            b.add_(torch.randn_like(b) * 0.05)
    actions = [str(i) for i in range(A)]
    rl = params.RLParameters(gamma=gamma, target_update_rate=tau, q_network_loss=loss,
                             maxq_learning=maxq, multi_steps=multi_steps,
                             use_seq_num_diff_as_time_diff=time_diff,
                             reward_boost=boost)
    trainer = tr.DQNTrainer(
        q, qt, None, actions=actions, rl=rl, double_q_learning=double_q, minibatch_size=B,
        optimizer=union.Optimizer__Union(Adam=union.classes["Adam"](lr=lr)),
        evaluation=params.EvaluationParameters(calc_cpe_in_training=False))
    act_idx = torch.randint(A, (B,))
    nact_idx = torch.randint(A, (B,))
    not_terminal = (torch.rand(B, 1) > 0.2).float()
    pnam = torch.ones(B, A)
    if random_masks:
        pnam = (torch.rand(B, A) > 0.3).float()
        pnam[torch.arange(B), torch.randint(A, (B,))] = 1.0
    batch = dict(
        state=torch.randn(B, S), next_state=torch.randn(B, S), reward=torch.randn(B, 1),
        time_diff=torch.randint(1, 4, (B, 1)).float(), step=torch.randint(1, 4, (B, 1)),
        not_terminal=not_terminal,
        action=torch.nn.functional.one_hot(act_idx, A).float(),
        next_action=torch.nn.functional.one_hot(nact_idx, A).float() * not_terminal,
        possible_actions_mask=torch.ones(B, A), possible_next_actions_mask=pnam)
    rbatch = rlt.DiscreteDqnInput(
        state=rlt.FeatureData(batch["state"]), next_state=rlt.FeatureData(batch["next_state"]),
        reward=batch["reward"], time_diff=batch["time_diff"],
        step=batch["step"] if multi_steps is not None else None,
        not_terminal=batch["not_terminal"], action=batch["action"],
        next_action=batch["next_action"], possible_actions_mask=batch["possible_actions_mask"],
        possible_next_actions_mask=batch["possible_next_actions_mask"],
        extras=rlt.ExtraData(action_probability=torch.ones(B, 1)))
    arrays = {f"batch.{k}": _np(v) for k, v in batch.items()}
    _dump_net(arrays, "q0", q)
    _dump_net(arrays, "qt0", qt)
    opts = [o["optimizer"] for o in trainer.configure_optimizers()]
    losses = []
    for it in range(N_UPDATES):
        cap = {}
        out = run_update(trainer, rbatch, it, opts, capture=cap)
        losses.append(out[0])
        if it == 0:
            for i, g in enumerate(cap[0]):
                arrays[f"grad0.{i}"] = _np(g)
            arrays["all_q0"] = _np(trainer.all_action_scores)
    arrays["losses"] = np.array(losses, dtype=np.float64)
    _dump_net(arrays, "qN", q)
    _dump_net(arrays, "qtN", qt)
    meta = dict(kind="dqn", B=B, S=S, A=A, sizes=list(sizes), acts=list(acts), loss=loss,
                double_q=double_q, maxq=maxq, multi_steps=multi_steps, time_diff=time_diff,
                boost=boost, gamma=gamma, tau=tau, lr=lr, n_updates=N_UPDATES)
    _save(name, arrays, meta)


def main():
    dqn_case("dqn_huber_double")
    dqn_case("dqn_mse_single_masked", loss="mse", double_q=False, random_masks=True, seed=1)
    dqn_case("dqn_sarsa", maxq=False, seed=2)
    dqn_case("dqn_multistep_boost", multi_steps=3, boost={"1": 0.5, "3": -0.25}, seed=3,
             acts=("leaky_relu", "tanh"))
    dqn_case("dqn_timediff_odd_dims", time_diff=True, B=37, S=7, A=3, sizes=(10, 6), seed=4)


if __name__ == "__main__":
    main()
