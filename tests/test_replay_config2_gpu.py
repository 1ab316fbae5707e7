"""BASELINE config-2 size replay parity: capacity 2^20 (depth-20 fp64 sum tree), B = 4096,
prioritized, bulk-filled -- the exact shape bench.py samples from.  Bit-exact indices,
gathered fields and sampling probabilities against the CPU oracle (oracle/replay_oracle.py,
pinned to the reference's buffers by tests/test_oracle_golden.py) on the same `random` seed,
including draws that take the reference's retry path (prioritized_replay_buffer.py:86-115)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

S, A, B, CAP = 128, 16, 4096, 1 << 20


def _stream(n, seed):
    rng = np.random.RandomState(seed)
    return dict(observation=rng.standard_normal((n, S)).astype(np.float32),
                action=rng.randint(0, A, n).astype(np.int64),
                reward=rng.standard_normal(n).astype(np.float32),
                terminal=rng.rand(n) < (1.0 / 200.0),
                priority=rng.uniform(0.1, 10.0, n))


@pytest.mark.parametrize("horizon", [1, 3])
def test_config2_size_per_sampling_matches_oracle(horizon):
    from oracle.replay_oracle import ReplayOracle
    from reagent_b200.replay_memory import PrioritizedReplayBuffer

    st = _stream(CAP, 1000)
    rb = PrioritizedReplayBuffer(stack_size=1, replay_capacity=CAP, batch_size=B,
                                 update_horizon=horizon, gamma=0.99)
    rb.add_batch(**st)
    ro = ReplayOracle(CAP, update_horizon=horizon, gamma=0.99, prioritized=True)
    ro.bulk_fill(st)
    assert np.array_equal(rb._is_index_valid.numpy(), ro.valid)
    # the fp64 heap (sequential delta propagation, sum_tree.py:164-189) bit for bit
    for lvl, ref in zip(rb.sum_tree.nodes, ro.tree.nodes):
        assert np.array_equal(np.asarray(lvl)[: len(ref)], ref)
    invalid = np.flatnonzero(~ro.valid)
    assert len(invalid) >= 1  # the newest `horizon` slots of the open episode
    # give the invalid slots real mass so that every draw takes the retry path a few times
    boost = np.full(len(invalid), 2.0e4)
    rb.set_priority(invalid.astype(np.int32), boost)
    for i, v in zip(invalid.tolist(), boost.tolist()):
        ro.tree.set(i, v)
    assert rb.sum_tree._total_priority() == ro.tree.total()

    for draw in range(3):
        random.seed(4242 + draw)
        got = rb.sample_discrete_dqn_batch(B, A)
        st_after = random.getstate()
        random.seed(4242 + draw)
        n_before = random.getstate()
        want = ro.sample_transition_batch(B)
        assert random.getstate() == st_after, "host random stream consumed differently"
        idx = got.indices.cpu().numpy().reshape(-1)
        assert np.array_equal(idx, want["indices"]), draw
        assert ro.valid[idx].all()
        assert np.array_equal(got.state.float_features.cpu().numpy(), want["state"])
        assert np.array_equal(got.next_state.float_features.cpu().numpy(), want["next_state"])
        onehot = np.eye(A, dtype=np.float32)[want["action"]]
        assert np.array_equal(got.action.cpu().numpy(), onehot)
        term = want["terminal"].astype(bool)
        n_onehot = np.eye(A, dtype=np.float32)[want["next_action"]] * (~term)[:, None]
        assert np.array_equal(got.next_action.cpu().numpy(), n_onehot)
        assert np.array_equal(got.not_terminal.cpu().numpy().reshape(-1), 1.0 - term.astype(np.float32))
        assert np.array_equal(got.step.cpu().numpy().reshape(-1), want["step"].astype(np.float32))
        if horizon == 1:
            assert np.array_equal(got.reward.cpu().numpy().reshape(-1), want["reward"])
        else:
            np.testing.assert_allclose(got.reward.cpu().numpy().reshape(-1), want["reward"],
                                       rtol=2e-6, atol=1e-6)
        assert np.array_equal(got.sampling_probabilities.cpu().numpy().reshape(-1),
                              want["sampling_probabilities"])
    # the retry path was really taken: with 2e4 of mass on each invalid slot (total ~5.3e6)
    # a 4096-strata draw lands on one ~15 times
    random.seed(4242)
    q, pos, idxs = rb.host_queries(B)
    assert len(pos) >= 1


def test_config2_size_uniform_sampling_matches_oracle():
    from oracle.replay_oracle import ReplayOracle
    from reagent_b200.replay_memory import ReplayBuffer

    st = _stream(CAP, 7)
    del st["priority"]
    rb = ReplayBuffer(stack_size=1, replay_capacity=CAP, batch_size=B)
    rb.add_batch(**st)
    ro = ReplayOracle(CAP, prioritized=False)
    ro.bulk_fill(st)
    assert np.array_equal(rb._is_index_valid.numpy(), ro.valid)
    for draw in range(2):
        torch.manual_seed(99 + draw)
        got = rb.sample_discrete_dqn_batch(B, A)
        torch.manual_seed(99 + draw)
        want = ro.sample_transition_batch(B)
        assert np.array_equal(got.indices.cpu().numpy().reshape(-1), want["indices"])
        assert np.array_equal(got.state.float_features.cpu().numpy(), want["state"])
        assert np.array_equal(got.next_state.float_features.cpu().numpy(), want["next_state"])
