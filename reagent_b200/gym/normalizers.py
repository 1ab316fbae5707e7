"""Default normalization maps for gym environments (reagent/gym/normalizers.py): every listed
feature gets the same type with mean 0 / stddev 1 and optional serving ranges."""
import collections

import numpy as np

from ..core.parameters import NormalizationParameters

_SCALAR_TYPES = (int, float, type(None))


def normalizer_helper(feats, feature_type, min_value=None, max_value=None):
    assert feature_type in ("DISCRETE_ACTION", "CONTINUOUS", "CONTINUOUS_ACTION"), \
        f"invalid feature type: {feature_type}."
    assert type(min_value) == type(max_value) and type(min_value) in _SCALAR_TYPES + (list, np.ndarray), \
        f"invalid {type(min_value)}, {type(max_value)}"
    if type(min_value) in _SCALAR_TYPES:
        min_value, max_value = [min_value] * len(feats), [max_value] * len(feats)
    opt = lambda v: None if v is None else float(v)  # noqa: E731
    return collections.OrderedDict(
        (f, NormalizationParameters(feature_type=feature_type, boxcox_lambda=None,
                                    boxcox_shift=None, mean=0, stddev=1, possible_values=None,
                                    quantiles=None, min_value=opt(lo), max_value=opt(hi)))
        for f, lo, hi in zip(feats, min_value, max_value))


def discrete_action_normalizer(feats):
    return normalizer_helper(feats, "DISCRETE_ACTION")


def only_continuous_normalizer(feats, min_value=None, max_value=None):
    return normalizer_helper(feats, "CONTINUOUS", min_value, max_value)


def only_continuous_action_normalizer(feats, min_value=None, max_value=None):
    return normalizer_helper(feats, "CONTINUOUS_ACTION", min_value, max_value)
