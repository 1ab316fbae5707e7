"""Feature type names, their order, and type identification from sampled values
(reagent/preprocessing/identify_types.py:9-73).  Host-side set-up code (numpy)."""
import numpy as np

BINARY = "BINARY"
PROBABILITY = "PROBABILITY"
CONTINUOUS = "CONTINUOUS"
BOXCOX = "BOXCOX"
ENUM = "ENUM"
QUANTILE = "QUANTILE"
CONTINUOUS_ACTION = "CONTINUOUS_ACTION"
DISCRETE_ACTION = "DISCRETE_ACTION"
DO_NOT_PREPROCESS = "DO_NOT_PREPROCESS"
CLIP_LOG = "CLIP_LOG"
FEATURE_TYPES = (BINARY, PROBABILITY, CONTINUOUS, BOXCOX, ENUM, QUANTILE, CONTINUOUS_ACTION,
                 DISCRETE_ACTION, DO_NOT_PREPROCESS, CLIP_LOG)

ROW_DELIM = "\n"
COLUMN_DELIM = ";"
DEFAULT_MAX_UNIQUE_ENUM = 10


def identify_type(values, enum_threshold=DEFAULT_MAX_UNIQUE_ENUM):
    """First matching rule of identify_types.py:63-73: BINARY (only 0/1, or a constant column),
    PROBABILITY (all within [0, 1]), ENUM (non-negative integers, at most `enum_threshold`
    distinct ones), else CONTINUOUS."""
    v = np.asarray(values)
    lo, hi = np.min(v), np.max(v)
    if lo == hi or bool(np.all((v == 0) | (v == 1))):
        return BINARY
    if lo >= 0 and hi <= 1:
        return PROBABILITY
    if lo >= 0 and len(np.unique(v)) <= enum_threshold and bool(np.all(np.floor(v) == v)):
        return ENUM
    return CONTINUOUS
