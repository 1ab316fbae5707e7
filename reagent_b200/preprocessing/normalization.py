"""Normalization parameters: constants, inference from sampled feature values and the helpers
around the dense preprocessor (reagent/preprocessing/normalization.py).  Host-side set-up code
(numpy / scipy), not on the per-step path; kept so that a workflow which infers its
NormalizationParameters with the reference's functions finds the same functions here."""
import json
from dataclasses import asdict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from ..core.parameters import NormalizationParameters
from . import identify_types
from .identify_types import DEFAULT_MAX_UNIQUE_ENUM, FEATURE_TYPES

BOX_COX_MAX_STDDEV = 1e8
BOX_COX_MARGIN = 1e-4
MISSING_VALUE = -1337.1337
DEFAULT_QUANTILE_K2_THRESHOLD = 1000.0
MINIMUM_SAMPLES_TO_IDENTIFY = 20
DEFAULT_MAX_QUANTILE_SIZE = 20
DEFAULT_NUM_SAMPLES = 100000
MAX_FEATURE_VALUE = 11.513  # logit of the clamped probability limits (1e-5, 1 - 1e-5)
MIN_FEATURE_VALUE = MAX_FEATURE_VALUE * -1
EPS = 1e-6


def no_op_feature():
    return NormalizationParameters(identify_types.CONTINUOUS, None, 0, 0, 1, None, None, None, None)


def _centre_scale(values):
    """(mean, stddev floored at 1, centred values) -- normalization.py:79-82, :149-156."""
    mean = float(np.mean(values))
    centred = values - mean
    return mean, max(float(np.std(centred, ddof=1)), 1.0), centred


def identify_parameter(feature_name, values, max_unique_enum_values=DEFAULT_MAX_UNIQUE_ENUM,
                       quantile_size=DEFAULT_MAX_QUANTILE_SIZE,
                       quantile_k2_threshold=DEFAULT_QUANTILE_K2_THRESHOLD, skip_box_cox=False,
                       skip_quantiles=False, feature_type=None):
    """NormalizationParameters of one feature from a sample of its values
    (normalization.py:45-173): type identification unless forced, Box-Cox when it makes the
    sample markedly more normal (D'Agostino K2), quantiles when the sample stays far from
    normal, mean / stddev for the standardised types, possible values for ENUM."""
    from scipy import stats
    from scipy.stats.mstats import mquantiles

    forced = feature_type
    if feature_type is None:
        feature_type = identify_types.identify_type(values, max_unique_enum_values)
    assert feature_type in FEATURE_TYPES, "unknown type {}".format(feature_type)
    assert len(values) >= MINIMUM_SAMPLES_TO_IDENTIFY, "insufficient information to identify parameter"
    force_boxcox = forced == identify_types.BOXCOX
    force_continuous = forced == identify_types.CONTINUOUS
    force_quantile = forced == identify_types.QUANTILE

    lam, shift, mean, stddev = None, 0.0, 0.0, 1.0
    possible_values = quantiles = None
    min_value, max_value = float(np.min(values)), float(np.max(values))

    if feature_type == identify_types.DO_NOT_PREPROCESS:
        mean, stddev, values = _centre_scale(values)

    if feature_type == identify_types.CONTINUOUS or force_boxcox or force_quantile:
        if max_value - min_value < BOX_COX_MARGIN and not (force_boxcox or force_quantile):
            return no_op_feature()
        k2_original, _ = stats.normaltest(values)
        shift = float(-min_value)  # (the shift could be estimated as well, scipy does not)
        candidate, lam_fit = stats.boxcox(np.maximum(values + shift, BOX_COX_MARGIN))
        k2_boxcox, _ = stats.normaltest(candidate)
        lambda_far_from_one = lam_fit < 0.9 or lam_fit > 1.1
        if (lambda_far_from_one or force_boxcox) and not (force_continuous or force_quantile):
            more_normal = k2_original > k2_boxcox * 10 and k2_boxcox <= quantile_k2_threshold
            if more_normal or force_boxcox:
                # (the reference stores this in `stddev` itself: it survives as the reported
                # stddev whenever the type ends up one that is not re-standardised below, e.g.
                # QUANTILE after skip_box_cox)
                stddev = cand_std = float(np.std(candidate, ddof=1))
                usable = np.isfinite(cand_std) and cand_std < BOX_COX_MAX_STDDEV and not np.isclose(cand_std, 0)
                if usable or force_boxcox:
                    values, lam = candidate, float(lam_fit)
        if lam is None or skip_box_cox:
            shift = lam = None
        if lam is not None:
            feature_type = identify_types.BOXCOX
        far_from_normal = lam is None and k2_original > quantile_k2_threshold
        if (far_from_normal and not skip_quantiles and not force_continuous) or force_quantile:
            feature_type = identify_types.QUANTILE
            probs = np.arange(quantile_size + 1, dtype=np.float64) / float(quantile_size)
            quantiles = np.unique(mquantiles(values, probs, alphap=0.0, betap=1.0)).astype(float).tolist()

    if feature_type in (identify_types.CONTINUOUS, identify_types.BOXCOX,
                        identify_types.CONTINUOUS_ACTION):
        mean, stddev, values = _centre_scale(values)
        if not np.isfinite(stddev):
            return None

    if feature_type == identify_types.ENUM:
        possible_values = np.unique(values.astype(int)).astype(int).tolist()

    return NormalizationParameters(feature_type, lam, shift, mean, stddev, possible_values,
                                   quantiles, min_value, max_value)


def get_feature_config(float_features: Optional[List[Tuple[int, str]]]):
    from ..core import types as rlt

    infos = [rlt.FloatFeatureInfo(name=name, feature_id=fid) for fid, name in (float_features or [])]
    return rlt.ModelFeatureConfig(float_feature_infos=infos)


def get_num_output_features(normalization_parameters: Dict[int, NormalizationParameters]) -> int:
    return sum(
        len(np_.possible_values) if np_.feature_type == identify_types.ENUM else 1
        for np_ in normalization_parameters.values())


def get_feature_start_indices(sorted_features: List[int],
                              normalization_parameters: Dict[int, NormalizationParameters]):
    """Start column of every feature in the preprocessor's output (ENUM fans out)."""
    starts, col = [], 0
    for f in sorted_features:
        p = normalization_parameters[f]
        starts.append(col)
        if p.feature_type == identify_types.ENUM:
            assert p.possible_values is not None
            col += len(p.possible_values)
        else:
            col += 1
    return starts


def sort_features_by_normalization(normalization_parameters: Dict[int, NormalizationParameters]
                                   ) -> Tuple[List[int], List[int]]:
    """(features ordered by FEATURE_TYPES then id, start of every type's section)."""
    assert isinstance(next(iter(normalization_parameters)), int), "Normalization Parameters need to be int"
    ordered: List[int] = []
    type_starts: List[int] = []
    ids = sorted(normalization_parameters)
    for ft in FEATURE_TYPES:
        type_starts.append(len(ordered))
        ordered += [f for f in ids if normalization_parameters[f].feature_type == ft]
    return ordered, type_starts


def serialize_one(feature_parameters):
    return json.dumps(asdict(feature_parameters))


def serialize(parameters):
    return {feature: serialize_one(p) for feature, p in parameters.items()}


def deserialize(parameters_json) -> Dict[int, NormalizationParameters]:
    out = {}
    for feature, blob in parameters_json.items():
        p = NormalizationParameters(**json.loads(blob))
        if p.feature_type == identify_types.ENUM:
            assert p.possible_values is not None
        out[int(feature)] = p
    return out


def get_feature_norm_metadata(feature_name, feature_value_list, norm_params):
    """identify_parameter with the options of a `norm_params` dict (normalization.py:264-298);
    None when fewer than MINIMUM_SAMPLES_TO_IDENTIFY values were collected."""
    if len(feature_value_list) < MINIMUM_SAMPLES_TO_IDENTIFY:
        return None
    override = None
    if norm_params["feature_overrides"] is not None:
        override = norm_params["feature_overrides"].get(feature_name, None)
    override = override or norm_params.get("default_feature_override", None)
    vals = np.array(feature_value_list, dtype=np.float32)
    assert not np.any(np.isinf(vals)), "Feature values contain infinity"
    assert not np.any(np.isnan(vals)), "Feature values contain nan (are there nulls in the feature values?)"
    return identify_parameter(feature_name, vals, norm_params["max_unique_enum_values"],
                              norm_params["quantile_size"], norm_params["quantile_k2_threshold"],
                              norm_params["skip_box_cox"], norm_params["skip_quantiles"], override)


def construct_action_scale_tensor(action_norm_params, action_scale_overrides):
    """(min, max) serving-range tensors [1, A] used to rescale continuous actions to [-1, 1]."""
    order, _ = sort_features_by_normalization(action_norm_params)
    lo = np.zeros((1, len(order)))
    hi = np.zeros((1, len(order)))
    for j, fid in enumerate(order):
        if fid in action_scale_overrides:
            lo[0][j], hi[0][j] = action_scale_overrides[fid][0], action_scale_overrides[fid][1]
        else:
            lo[0][j], hi[0][j] = action_norm_params[fid].min_value, action_norm_params[fid].max_value
    return torch.from_numpy(lo), torch.from_numpy(hi)
