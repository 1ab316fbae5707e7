"""Offline batch formatters against vectors produced by the UNMODIFIED
reference classes (oracle/make_golden.py: batch_preprocessor_case)."""
import numpy as np
import pytest
import torch

from tests import golden_util as G


def _batch(arrays, prefix="in.", extra=()):
    b = {k[len(prefix):]: torch.from_numpy(np.asarray(v)) for k, v in arrays.items()
         if k.startswith(prefix)}
    for k in extra:
        b[k] = torch.from_numpy(np.asarray(arrays["cin." + k]))
    return b


class _Identity:
    """Stands in for the CUDA Preprocessor in the CPU test of the index arithmetic."""
    device = torch.device("cpu")

    def __call__(self, x, presence):
        return x


def test_discrete_batch_fields_cpu():
    from reagent_b200.preprocessing import DiscreteDqnBatchPreprocessor

    arrays, meta = G.load("batch_preprocessor")
    out = DiscreteDqnBatchPreprocessor(meta["A"], _Identity())(_batch(arrays))
    # bit-exact: integer / index work and reshapes
    for name, got in [("action", out.action), ("next_action", out.next_action),
                      ("reward", out.reward), ("time_diff", out.time_diff), ("step", out.step),
                      ("not_terminal", out.not_terminal), ("mdp_id", out.extras.mdp_id),
                      ("sequence_number", out.extras.sequence_number),
                      ("action_probability", out.extras.action_probability)]:
        ref = arrays["d." + name]
        assert tuple(got.shape) == ref.shape, name
        assert np.array_equal(got.numpy(), ref), name
    assert out.next_action.shape[1] == meta["A"]
    assert float(out.not_terminal[:5].sum()) == 0.0  # rows without a possible next action


def _preprocessors(meta):
    from reagent_b200.core.parameters import NormalizationParameters as NP
    from reagent_b200.preprocessing import Preprocessor

    sp = Preprocessor({int(k): NP(**v) for k, v in meta["s_spec"].items()}).eval()
    ap = Preprocessor({int(k): NP(**v) for k, v in meta["a_spec"].items()}).eval()
    return sp, ap


@pytest.mark.gpu
def test_discrete_batch_matches_reference_gpu():
    from reagent_b200.preprocessing import DiscreteDqnBatchPreprocessor

    arrays, meta = G.load("batch_preprocessor")
    sp, _ = _preprocessors(meta)
    out = DiscreteDqnBatchPreprocessor(meta["A"], sp, use_gpu=True)(_batch(arrays))
    assert out.state.float_features.is_cuda
    assert G.rel_err(out.state.float_features, arrays["d.state"]) < 1e-6
    assert G.rel_err(out.next_state.float_features, arrays["d.next_state"]) < 1e-6
    for name, got in [("action", out.action), ("next_action", out.next_action),
                      ("not_terminal", out.not_terminal), ("step", out.step)]:
        assert np.array_equal(got.cpu().numpy(), arrays["d." + name]), name


@pytest.mark.gpu
def test_policy_batch_matches_reference_gpu():
    from reagent_b200.preprocessing import PolicyNetworkBatchPreprocessor

    arrays, meta = G.load("batch_preprocessor")
    sp, ap = _preprocessors(meta)
    b = _batch(arrays)
    b.update(_batch(arrays, prefix="cin."))
    out = PolicyNetworkBatchPreprocessor(sp, ap, use_gpu=True)(b)
    assert G.rel_err(out.state.float_features, arrays["c.state"]) < 1e-6
    assert G.rel_err(out.action.float_features, arrays["c.action"]) < 1e-6
    assert G.rel_err(out.next_action.float_features, arrays["c.next_action"]) < 1e-6
    assert np.array_equal(out.not_terminal.cpu().numpy(), arrays["c.not_terminal"])
    assert np.array_equal(out.reward.cpu().numpy(), arrays["c.reward"])
