"""Pins the CPU restatement (oracle/td_oracle.py) against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import td_oracle as O
from tests import golden_util as G

DQN_CASES = ["dqn_huber_double", "dqn_mse_single_masked", "dqn_sarsa", "dqn_multistep_boost",
             "dqn_timediff_odd_dims"]


def _dqn_kwargs(meta, batch):
    kw = dict(double_q=meta["double_q"], maxq=meta["maxq"], loss=meta["loss"])
    if meta["multi_steps"] is not None:
        kw["discount_src"] = batch["step"]
    elif meta["time_diff"]:
        kw["discount_src"] = batch["time_diff"]
    if meta["boost"]:
        rb = torch.zeros(1, meta["A"])
        for k, v in meta["boost"].items():
            rb[0, int(k)] = v
        kw["reward_boost"] = rb
    return kw


@pytest.mark.parametrize("name", DQN_CASES)
def test_dqn_oracle_matches_reference(name):
    arrays, meta = G.load(name)
    acts = meta["acts"] + ["linear"]
    q = G.oracle_net(arrays, "q0", acts, requires_grad=True)
    qt = G.oracle_net(arrays, "qt0", acts)
    batch = G.batch_tensors(arrays)
    adam = O.AdamState(O.net_params(q), lr=meta["lr"])
    kw = _dqn_kwargs(meta, batch)
    for it in range(meta["n_updates"]):
        loss, grads, aux = O.dqn_update(q, qt, adam, batch, gamma=meta["gamma"], tau=meta["tau"], **kw)
        assert abs(loss - arrays["losses"][it]) <= 1e-6 * max(1.0, abs(arrays["losses"][it]))
        if it == 0:
            for i, g in enumerate(grads):
                assert G.rel_err(g, arrays[f"grad0.{i}"]) < 1e-6
            assert G.rel_err(aux["all_q"], arrays["all_q0"]) < 1e-6
    for i in range(len(q["W"])):
        assert G.rel_err(q["W"][i], arrays[f"qN.W{i}"]) < 1e-6
        assert G.rel_err(q["b"][i], arrays[f"qN.b{i}"]) < 1e-6
        assert G.rel_err(qt["W"][i], arrays[f"qtN.W{i}"]) < 1e-6
        assert G.rel_err(qt["b"][i], arrays[f"qtN.b{i}"]) < 1e-6


# ---------------------------------------------------------------------------
# replay sampler restatement vs the reference buffers
# ---------------------------------------------------------------------------
REPLAY_CASES = ["replay_uniform_h1", "replay_uniform_h3_wrap", "replay_uniform_h5_cont",
                "replay_per_h1", "replay_per_h3_wrap_zero", "replay_per_big"]


@pytest.mark.parametrize("name", REPLAY_CASES)
def test_replay_oracle_matches_reference(name):
    import random

    import numpy as np

    from oracle.replay_oracle import ReplayOracle

    arrays, meta = G.load(name)
    rb = ReplayOracle(meta["cap"], meta["horizon"], meta["gamma"], meta["prioritized"])
    keys = meta["keys"]
    for t in range(meta["n_add"]):
        rb.add(**{k: arrays[f"stream.{k}"][t] for k in keys})
    assert np.array_equal(rb.valid, arrays["valid"])
    random.seed(meta["seed"] + 100)
    torch.manual_seed(meta["seed"] + 100)
    for s_i in range(meta["n_samples"]):
        out = rb.sample_transition_batch(meta["B"])
        for f, v in out.items():
            key = f"sample{s_i}.{f}"
            if key not in arrays:
                continue
            want = arrays[key]
            got = np.asarray(v).reshape(want.shape)
            if f.startswith("next_"):
                # "When the transition is terminal next_state_batch has undefined contents"
                # (circular_replay_buffer.py:621): the reference may read np.empty() memory.
                keep = ~arrays[f"sample{s_i}.terminal"].reshape(-1)
                got, want = got[keep], want[keep]
            assert np.array_equal(got, want), key
