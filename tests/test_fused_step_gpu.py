"""FusedDqnStep (whole update as one CUDA graph) against the eager path it captures:
`rb.sample_discrete_dqn_batch` + `trainer.train_batch` with the same host random stream.
Same kernels on the same data, so the loss sequences must agree bit for bit."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

S, A, B, CAP = 24, 5, 256, 4096


def _stream(n, seed):
    rng = np.random.RandomState(seed)
    return dict(observation=rng.standard_normal((n, S)).astype(np.float32),
                action=rng.randint(0, A, n).astype(np.int64),
                reward=rng.standard_normal(n).astype(np.float32),
                terminal=rng.rand(n) < 0.02, priority=rng.uniform(0.1, 10.0, n))


def _setup(prioritized):
    from reagent_b200.core.parameters import EvaluationParameters, RLParameters
    from reagent_b200.models import FullyConnectedDQN
    from reagent_b200.optimizer import Optimizer__Union
    from reagent_b200.replay_memory import PrioritizedReplayBuffer, ReplayBuffer
    from reagent_b200.training import DQNTrainer

    dev = torch.device("cuda", 0)
    data = _stream(CAP - 7, 3)
    if prioritized:
        rb = PrioritizedReplayBuffer(stack_size=1, replay_capacity=CAP, batch_size=B, device=dev)
    else:
        rb = ReplayBuffer(stack_size=1, replay_capacity=CAP, batch_size=B, device=dev)
        data.pop("priority")
    rb.add_batch(**data)
    torch.manual_seed(1)
    q = FullyConnectedDQN(S, A, [48, 32], ["relu", "relu"])
    qt = q.get_target_network()
    t = DQNTrainer(q.to(dev), qt.to(dev), actions=[str(i) for i in range(A)],
                   rl=RLParameters(gamma=0.9, target_update_rate=0.05, q_network_loss="huber"),
                   double_q_learning=True, minibatch_size=B,
                   optimizer=Optimizer__Union.default(lr=1e-2),
                   evaluation=EvaluationParameters(calc_cpe_in_training=False)).to(dev)
    return rb, t


def _seed():
    random.seed(7)
    torch.manual_seed(7)
    np.random.seed(7)


@pytest.mark.parametrize("prioritized", [True, False])
@pytest.mark.parametrize("prefetch", [False, True])
def test_fused_step_matches_eager(prioritized, prefetch):
    from reagent_b200.training.fused_step import FusedDqnStep

    n = 7
    rb, t = _setup(prioritized)
    _seed()
    eager = []
    for _ in range(n + 1):  # FusedDqnStep's constructor runs one warm-up update
        eager.append(float(t.train_batch(rb.sample_discrete_dqn_batch(B, A))))
    rb2, t2 = _setup(prioritized)
    _seed()
    fused = FusedDqnStep(t2, rb2, B, prefetch=prefetch)
    got = []
    for _ in range(n):
        lh = fused.step()
        torch.cuda.synchronize()
        got.append(float(lh[0]))
    assert got == eager[1:], (got, eager[1:])
    for a, b in zip(t.q_network.parameters(), t2.q_network.parameters()):
        assert torch.equal(a, b)
    for a, b in zip(t.q_network_target.parameters(), t2.q_network_target.parameters()):
        assert torch.equal(a, b)


def test_fused_step_sees_parameters_loaded_from_outside():
    """load_state_dict between two replays of the captured update: the tensor-core weight
    images of K2 are rebuilt before the next replay (FusedDqnStep._refresh_tc_images), so the
    captured path keeps matching the eager one bit for bit."""
    from reagent_b200.models import FullyConnectedDQN
    from reagent_b200.training.fused_step import FusedDqnStep

    torch.manual_seed(99)
    donor = FullyConnectedDQN(S, A, [48, 32], ["relu", "relu"]).state_dict()
    rb, t = _setup(True)
    _seed()
    eager = []
    for i in range(6):
        if i == 3:
            t.q_network.load_state_dict(donor)
        eager.append(float(t.train_batch(rb.sample_discrete_dqn_batch(B, A))))
    rb2, t2 = _setup(True)
    _seed()
    fused = FusedDqnStep(t2, rb2, B, prefetch=False)
    got = []
    for i in range(1, 6):
        if i == 3:
            t2.q_network.load_state_dict(donor)
        lh = fused.step()
        torch.cuda.synchronize()
        got.append(float(lh[0]))
    assert got == eager[1:], (got, eager[1:])
    for a, b in zip(t.q_network.parameters(), t2.q_network.parameters()):
        assert torch.equal(a, b)
