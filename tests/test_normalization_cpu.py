"""P2: normalization parameter inference against the unmodified reference
(reagent/preprocessing/normalization.py:45-173, identify_types.py:63-73) on golden samples
(oracle/make_golden.py::normalization_case), plus the gym normalizers and the helpers."""
import dataclasses
import math

import numpy as np
import pytest

from tests import golden_util as G


def _same(a, b, path=""):
    if isinstance(a, float) and isinstance(b, float):
        assert a == b or (math.isnan(a) and math.isnan(b)) or abs(a - b) <= 1e-12 * max(1.0, abs(b)), (path, a, b)
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    else:
        assert a == b, (path, a, b)


def test_identify_parameter_matches_reference():
    from reagent_b200.preprocessing import identify_types, normalization

    arrays, meta = G.load("normalization_identify")
    exp = meta["expected"]
    for key, want in exp.items():
        parts = key.split(".")
        v = arrays[f"values.{parts[1]}"].copy()
        if parts[0] == "auto":
            assert identify_types.identify_type(v) == want["type"], key
            got, want = normalization.identify_parameter(parts[1], v), want["params"]
        elif parts[0] == "forced":
            got = normalization.identify_parameter(parts[1], v, feature_type=parts[2])
        elif parts[0] == "skip_box_cox":
            got = normalization.identify_parameter(parts[1], v, skip_box_cox=True)
        else:
            got = normalization.identify_parameter(parts[1], v, skip_quantiles=True)
        if want is None:
            assert got is None, key
            continue
        got = dataclasses.asdict(got)
        assert set(got) == set(want), key
        for f in want:
            _same(got[f], want[f], f"{key}.{f}")


def test_normalization_helpers_and_gym_normalizers():
    from reagent_b200.core.parameters import NormalizationParameters as NP
    from reagent_b200.gym import normalizers
    from reagent_b200.preprocessing import normalization as N

    params = {3: NP("ENUM", possible_values=[1, 4, 9]), 1: NP("CONTINUOUS", mean=1.0, stddev=2.0),
              2: NP("BINARY"), 7: NP("CONTINUOUS")}
    order, starts = N.sort_features_by_normalization(params)
    assert order == [2, 1, 7, 3]
    assert starts == [0, 1, 1, 3, 3, 4, 4, 4, 4, 4]
    assert N.get_feature_start_indices(order, params) == [0, 1, 2, 3]
    assert N.get_num_output_features(params) == 6
    back = N.deserialize(N.serialize(params))
    assert back == params
    cfg = N.get_feature_config([(5, "a"), (6, "b")])
    assert [f.feature_id for f in cfg.float_feature_infos] == [5, 6] and cfg.only_dense
    m = normalizers.only_continuous_normalizer([0, 1, 2], -1.0, 2.0)
    assert list(m) == [0, 1, 2] and m[1].feature_type == "CONTINUOUS" and m[1].min_value == -1.0
    assert m[2].max_value == 2.0 and m[0].mean == 0 and m[0].stddev == 1
    a = normalizers.only_continuous_action_normalizer([4, 5], [-1, -2], [1, 2])
    assert a[5].min_value == -2.0 and a[5].feature_type == "CONTINUOUS_ACTION"
    assert normalizers.discrete_action_normalizer([9])[9].feature_type == "DISCRETE_ACTION"
    lo, hi = N.construct_action_scale_tensor(a, {4: (-3.0, 3.0)})
    assert lo.tolist() == [[-3.0, -2.0]] and hi.tolist() == [[3.0, 2.0]]
    with pytest.raises(AssertionError):
        N.identify_parameter("x", np.arange(5, dtype=np.float32))
    md = N.get_feature_norm_metadata("x", list(np.linspace(0, 50, 100)), dict(
        feature_overrides=None, max_unique_enum_values=10, quantile_size=20,
        quantile_k2_threshold=1000.0, skip_box_cox=False, skip_quantiles=False))
    assert md.feature_type == "CONTINUOUS"
