"""GPU parity of the FUSED trainer-batch samplers -- the path bench.py times -- against the
reference's `sample_transition_batch` followed by DiscreteDqnInputMaker / PolicyNetworkInputMaker
(reagent/gym/preprocessors/trainer_preprocessor.py:100-227, reagent/training/utils.py:13-29).
Golden vectors: oracle/make_golden.py::inputmaker_case (unmodified reference classes)."""
import random

import numpy as np
import pytest
import torch

from tests import golden_util as G

pytestmark = pytest.mark.gpu

CASES = ["inputmaker_dqn_uniform", "inputmaker_dqn_per_masks", "inputmaker_policy_uniform",
         "inputmaker_policy_per_h3"]


def _build(arrays, meta, bulk):
    from reagent_b200.replay_memory import PrioritizedReplayBuffer, ReplayBuffer

    cls = PrioritizedReplayBuffer if meta["prioritized"] else ReplayBuffer
    rb = cls(stack_size=1, replay_capacity=meta["cap"], batch_size=meta["B"],
             update_horizon=meta["horizon"], gamma=meta["gamma"])
    st = {k: arrays[f"stream.{k}"] for k in meta["keys"]}
    if bulk:
        rb.add_batch(**st)
        return rb
    for t in range(meta["n_add"]):
        kw = {}
        for k in meta["keys"]:
            v = st[k][t]
            if k == "terminal":
                v = bool(v)
            elif k == "priority":
                v = float(v)
            elif k == "action" and not meta["continuous"]:
                v = int(v)
            elif np.ndim(v) == 0:
                v = float(v)
            kw[k] = v
        rb.add(**kw)
    return rb


def _eq(name, got, want, keep=None):
    got = got.detach().cpu().numpy()
    assert got.shape == want.shape and got.dtype == want.dtype, (name, got.shape, want.shape, got.dtype, want.dtype)
    if keep is not None:
        got, want = got[keep], want[keep]
    assert np.array_equal(got, want), name


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("bulk", [False, True])
def test_fused_batch_matches_reference_inputmaker(name, bulk):
    arrays, meta = G.load(name)
    rb = _build(arrays, meta, bulk)
    B, A = meta["B"], meta["A"]
    random.seed(meta["seed"] + 200)
    torch.manual_seed(meta["seed"] + 200)
    np.random.seed(meta["seed"] + 200)
    for s_i in range(meta["n_samples"]):
        if meta["continuous"]:
            out = rb.sample_policy_network_batch(B, arrays["action_low"], arrays["action_high"])
            action, next_action = out.action.float_features, out.next_action.float_features
        else:
            out = rb.sample_discrete_dqn_batch(B, A)
            action, next_action = out.action, out.next_action
        pre = f"sample{s_i}."
        _eq(pre + "indices", out.indices, arrays[pre + "indices"])
        # "When the transition is terminal next_state_batch has undefined contents"
        # (circular_replay_buffer.py:621): compared on non-terminal rows only
        nonterm = ~arrays[pre + "terminal"].reshape(-1)
        _eq(pre + "state", out.state.float_features, arrays[pre + "state"])
        _eq(pre + "next_state", out.next_state.float_features, arrays[pre + "next_state"], nonterm)
        _eq(pre + "not_terminal", out.not_terminal, arrays[pre + "not_terminal"])
        _eq(pre + "action", action, arrays[pre + "action"])
        _eq(pre + "next_action", next_action, arrays[pre + "next_action"])  # zeroed on terminal rows
        # n-step fold: same fp32 products; torch.sum's order over the horizon may differ
        np.testing.assert_allclose(out.reward.cpu().numpy(), arrays[pre + "reward"],
                                   rtol=2e-6, atol=1e-6)
        if meta["horizon"] == 1:
            _eq(pre + "reward", out.reward, arrays[pre + "reward"])
        # exp() on the device vs the host: 1 ulp
        np.testing.assert_allclose(out.extras.action_probability.cpu().numpy(),
                                   arrays[pre + "action_probability"], rtol=3e-7, atol=0)
        if not meta["continuous"]:
            _eq(pre + "possible_actions_mask", out.possible_actions_mask,
                arrays[pre + "possible_actions_mask"])
            _eq(pre + "possible_next_actions_mask", out.possible_next_actions_mask,
                arrays[pre + "possible_next_actions_mask"], nonterm)
        assert out.time_diff is None
