"""Act-time samplers / scorer against vectors from the UNMODIFIED reference
(oracle/make_golden.py: sampler_case).  The samplers are device-agnostic score arithmetic with
torch's RNG, so seeded CPU draws reproduce the reference's draws bit for bit."""
import numpy as np
import pytest
import torch

from tests import golden_util as G


def _load():
    arrays, meta = G.load("act_samplers")
    t = {k: torch.from_numpy(np.asarray(v)) for k, v in arrays.items()}
    return t, meta


def test_greedy_sampler():
    from reagent_b200.gym.policies import GreedyActionSampler

    t, meta = _load()
    s = GreedyActionSampler()
    out = s.sample_action(t["scores"])
    assert torch.equal(out.action, t["greedy.action"]) and torch.equal(out.log_prob, t["greedy.log_prob"])
    assert torch.equal(s.log_prob(t["scores"], t["action"]), t["greedy.lp_of_action"])


def test_epsilon_greedy_sampler_reproduces_seeded_draws():
    from reagent_b200.gym.policies import EpsilonGreedyActionSampler

    t, meta = _load()
    s = EpsilonGreedyActionSampler(epsilon=0.3, epsilon_decay=0.5, minimum_epsilon=0.1)
    p = s.action_probabilities(t["masked"])
    assert float(p[t["masked"] <= -1e10].abs().max()) == 0.0  # invalid actions are never drawn
    torch.manual_seed(meta["seed"] + 1)
    out = s.sample_action(t["masked"])
    assert torch.equal(out.action, t["eps.action"]) and torch.equal(out.log_prob, t["eps.log_prob"])
    torch.manual_seed(meta["seed"] + 2)
    assert torch.equal(s.log_prob(t["masked"], t["action"]), t["eps.lp_of_action"])
    s.update(); s.update(); s.update()
    assert s.epsilon == float(t["eps.epsilon_after_3_updates"][0])


def test_softmax_sampler_reproduces_seeded_draws():
    from reagent_b200.gym.policies import SoftmaxActionSampler

    t, meta = _load()
    s = SoftmaxActionSampler(temperature=0.7, temperature_decay=0.5, minimum_temperature=0.2)
    torch.manual_seed(meta["seed"] + 3)
    out = s.sample_action(t["scores"])
    assert torch.equal(out.action, t["soft.action"]) and torch.equal(out.log_prob, t["soft.log_prob"])
    assert torch.equal(s.log_prob(t["scores"], t["action"]), t["soft.lp_of_action"])
    assert torch.equal(s.entropy(t["scores"]), t["soft.entropy"])
    s.update(); s.update()
    assert s.temperature == float(t["soft.temperature_after_2_updates"][0])
    with pytest.raises(AssertionError):
        SoftmaxActionSampler(temperature=0.0)


def test_possible_actions_mask_and_policy_composition():
    from reagent_b200.core import types as rlt
    from reagent_b200.gym.policies import GreedyActionSampler, Policy, apply_possible_actions_mask

    t, meta = _load()
    got = apply_possible_actions_mask(t["scores"][:1].clone(), t["mask_one"])
    assert torch.equal(got, t["masked_one"])
    calls = []

    def scorer(obs, mask=None):
        calls.append(mask)
        return apply_possible_actions_mask(t["scores"][:1].clone(), mask)

    pol = Policy(scorer=scorer, sampler=GreedyActionSampler())
    a = pol.act(rlt.FeatureData(torch.zeros(1, 3)), t["mask_one"])
    assert a.action.device.type == "cpu" and int(a.action.argmax()) == int(t["masked_one"].argmax())
    pol.act(rlt.FeatureData(torch.zeros(1, 3)))
    assert calls[0] is t["mask_one"] and calls[1] is None


@pytest.mark.gpu
def test_discrete_dqn_scorer_on_the_fused_forward():
    """scores = q_network(obs) through the fused forward kernel (plain, dueling and QR heads)."""
    from reagent_b200.core import types as rlt
    from reagent_b200.gym.policies import GreedyActionSampler, Policy, discrete_dqn_scorer
    from reagent_b200.models import DuelingQNetwork, FullyConnectedDQN

    torch.manual_seed(0)
    x = torch.randn(5, 12)
    for q in (FullyConnectedDQN(12, 4, [16, 8], ["relu", "relu"]),
              DuelingQNetwork.make_fully_connected(12, 4, [16, 8], ["relu", "relu"]),
              FullyConnectedDQN(12, 4, [16], ["relu"], num_atoms=7)):
        ref_params = [p.detach().clone() for p in q.parameters()]
        q = q.cuda()
        scorer = discrete_dqn_scorer(q)
        scores = scorer(rlt.FeatureData(x.cuda()))
        assert scores.shape == (5, 4) and q.training
        expect = q(rlt.FeatureData(x.cuda()))
        if expect.dim() == 3:
            expect = expect.mean(dim=2)
        assert torch.allclose(scores, expect)
        mask = torch.tensor([True, False, True, True])
        one = scorer(rlt.FeatureData(x[:1].cuda()), mask)
        assert float(one[0, 1]) == float("-inf")
        act = Policy(scorer, GreedyActionSampler()).act(rlt.FeatureData(x[:1].cuda()), mask)
        assert act.action.device.type == "cpu" and int(act.action.argmax()) != 1
        for p, r in zip(q.parameters(), ref_params):
            assert torch.equal(p.detach().cpu(), r)  # acting does not touch the parameters
