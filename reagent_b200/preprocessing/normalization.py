"""Constants and helpers of reagent/preprocessing/normalization.py that the hot path uses
(:25-45, :188-198).  Parameter inference (`identify_parameter`, scipy) is host-side setup
and out of scope; construct NormalizationParameters directly."""
from typing import Dict

from ..core.parameters import NormalizationParameters
from . import identify_types

MISSING_VALUE = -1337.1337
MAX_FEATURE_VALUE = 11.513
MIN_FEATURE_VALUE = MAX_FEATURE_VALUE * -1
EPS = 1e-6


def no_op_feature():
    return NormalizationParameters(identify_types.CONTINUOUS, None, 0, 0, 1, None, None, None, None)


def get_num_output_features(normalization_parameters: Dict[int, NormalizationParameters]) -> int:
    return sum(
        len(np.possible_values) if np.feature_type == identify_types.ENUM else 1
        for np in normalization_parameters.values())


def only_continuous_normalizer(feats, mean=0.0, stddev=1.0):
    """reagent/gym/normalizers.py: every feature CONTINUOUS with the given mean/stddev."""
    return {f: NormalizationParameters(identify_types.CONTINUOUS, mean=mean, stddev=stddev)
            for f in feats}
