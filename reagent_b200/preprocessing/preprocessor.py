"""Preprocessor with the reference's constructor / forward contract
(reagent/preprocessing/preprocessor.py:23-170): `forward(input (B,F) f32,
input_presence_byte (B,F)) -> (B,F') f32`, features sorted by FEATURE_TYPES order then id,
ENUM features expanded to one column per possible value.  One CUDA launch
(rb200_preprocess); the same column program can be fused into the replay gather
(ReplayBuffer.set_state_preprocessor)."""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch.nn import Module

from .. import _lib
from ..core.parameters import NormalizationParameters
from .identify_types import DO_NOT_PREPROCESS, ENUM, FEATURE_TYPES
from .normalization import EPS, MAX_FEATURE_VALUE, MIN_FEATURE_VALUE

_COL_DTYPE = np.dtype([("src_col", "<i4"), ("type", "<i4"), ("p0", "<f4"), ("p1", "<f4"),
                       ("p2", "<f4"), ("p3", "<f4"), ("q_off", "<i4"), ("q_cnt", "<i4")])
_RANGE_UNCHECKED = ("BOXCOX", "CONTINUOUS", "DO_NOT_PREPROCESS", "CLIP_LOG")


def _f32(x) -> float:
    return float(torch.tensor(x, dtype=torch.float32))


class Preprocessor(Module):
    def __init__(
        self,
        normalization_parameters: Dict[int, NormalizationParameters],
        use_gpu: Optional[bool] = None,
        device: Optional[torch.device] = None,
    ) -> None:
        super().__init__()
        self.normalization_parameters = normalization_parameters
        (self.feature_id_to_index, self.sorted_features,
         self.sorted_feature_boundaries) = self._sort_features_by_normalization()
        if device is not None:
            self.device = torch.device(device)
        elif torch.cuda.is_available():
            self.device = torch.device("cuda")
        else:
            self.device = torch.device("cpu")  # construction works; forward needs CUDA
        self._build_program()
        self._dev_prog = None

    # ---- reference helpers ---------------------------------------------------
    def _sort_features_by_normalization(self):
        """preprocessor.py:527-545"""
        feature_id_to_index = {}
        sorted_features = []
        feature_starts = []
        assert isinstance(list(self.normalization_parameters.keys())[0], int), (
            "Normalization Parameters need to be int")
        for feature_type in FEATURE_TYPES:
            feature_starts.append(len(sorted_features))
            for feature in sorted(self.normalization_parameters.keys()):
                norm = self.normalization_parameters[feature]
                if norm.feature_type == feature_type:
                    feature_id_to_index[feature] = len(sorted_features)
                    sorted_features.append(feature)
        return feature_id_to_index, sorted_features, feature_starts

    def input_prototype(self) -> Tuple[torch.Tensor, torch.Tensor]:
        n = len(self.normalization_parameters)
        return (torch.randn(1, n, device=self.device),
                torch.ones(1, n, dtype=torch.uint8, device=self.device))

    # ---- column program ------------------------------------------------------
    def _build_program(self):
        cols = []
        quant: List[float] = []
        self._sections = []  # (feature_type, out_begin, out_end) for the range check
        for ftype in FEATURE_TYPES:
            feats = [(i, f) for i, f in enumerate(self.sorted_features)
                     if self.normalization_parameters[f].feature_type == ftype]
            if not feats:
                continue
            begin = len(cols)
            tid = FEATURE_TYPES.index(ftype)
            if ftype == "QUANTILE":
                bmax = max(len(self.normalization_parameters[f].quantiles) for _, f in feats)
            for src, f in feats:
                p = self.normalization_parameters[f]
                c = dict(src_col=src, type=tid, p0=0.0, p1=0.0, p2=0.0, p3=0.0, q_off=0, q_cnt=0)
                if ftype == "PROBABILITY":
                    c.update(p0=_f32(1e-5), p1=_f32(1 - 1e-5))
                elif ftype == "CONTINUOUS":
                    c.update(p0=_f32(p.mean), p1=_f32(p.stddev))
                elif ftype == "BOXCOX":
                    assert abs(p.boxcox_lambda) > 1e-6, (
                        "Invalid value for boxcox lambda: " + str(p.boxcox_lambda))
                    c.update(p0=_f32(p.mean), p1=_f32(p.stddev), p2=_f32(p.boxcox_shift),
                             p3=_f32(p.boxcox_lambda))
                elif ftype == "ENUM":
                    for v in p.possible_values:
                        cv = dict(c)
                        cv["p0"] = _f32(float(v))
                        cols.append(cv)
                    continue
                elif ftype == "QUANTILE":
                    qs = list(p.quantiles) + [p.quantiles[-1]] * (bmax - len(p.quantiles))
                    c.update(p0=float(len(p.quantiles)) - 1, p1=_f32(max(p.quantiles)),
                             p2=_f32(min(p.quantiles)), q_off=len(quant), q_cnt=bmax)
                    quant += [_f32(q) for q in qs]
                elif ftype == "CONTINUOUS_ACTION":
                    # parameters built exactly as preprocessor.py:248-272 (fp32 tensor math)
                    scaling = ((torch.ones(1) - EPS) * 2
                               / torch.tensor([p.max_value - p.min_value]))
                    min_training = torch.ones(1) * -1 + EPS
                    c.update(p0=_f32(p.min_value), p1=float(scaling[0]),
                             p2=float(min_training[0]), p3=_f32(-1 + EPS))
                cols.append(c)
            self._sections.append((ftype, begin, len(cols)))
        arr = np.zeros(len(cols), dtype=_COL_DTYPE)
        for i, c in enumerate(cols):
            for k, v in c.items():
                arr[i][k] = v
        self._cols_host = arr
        self._quant_host = np.asarray(quant if quant else [0.0], dtype=np.float32)
        self.num_output_features = len(cols)

    def device_program(self, device):
        """(cols, quantiles, F') tensors on `device` for the CUDA kernels."""
        device = torch.device(device)
        if self._dev_prog is None or self._dev_prog[0].device != device:
            cols = torch.from_numpy(self._cols_host.view(np.uint8).copy()).to(device)
            quant = torch.from_numpy(self._quant_host.copy()).to(device)
            self._dev_prog = (cols, quant)
        return self._dev_prog[0], self._dev_prog[1], self.num_output_features

    # ---- forward ---------------------------------------------------------------
    def forward(self, input: torch.Tensor, input_presence_byte: torch.Tensor) -> torch.Tensor:
        assert input.shape == input_presence_byte.shape, (
            f"{input.shape} != {input_presence_byte.shape}")
        if not input.is_cuda:
            raise _lib.Rb200Error("Preprocessor.forward: reagent_b200 runs on CUDA only "
                                  f"(got a {input.device} tensor); there is no CPU fallback")
        x = input.contiguous().float()
        pres = input_presence_byte.contiguous()
        if pres.dtype == torch.float32:
            is_float = 1
        elif pres.dtype in (torch.uint8, torch.bool):
            is_float = 0
        else:
            pres = pres.float()
            is_float = 1
        rows, f_in = x.shape
        cols, quant, f_out = self.device_program(x.device)
        out = torch.empty(rows, f_out, dtype=torch.float32, device=x.device)
        rc = _lib.lib().rb200_preprocess(x.data_ptr(), pres.data_ptr(), is_float, rows, f_in,
                                         f_out, cols.data_ptr(), quant.data_ptr(),
                                         out.data_ptr(), _lib.cur_stream())
        _lib.check(rc, "rb200_preprocess")
        if self.training:
            self._check_preprocessing_output(out)
        return out

    def _check_preprocessing_output(self, out):
        """preprocessor.py:576-599 (host-syncing range check in training mode)."""
        for ftype, b, e in self._sections:
            if ftype in _RANGE_UNCHECKED or out.shape[0] == 0:
                continue
            sec = out[:, b:e]
            max_value, min_value = sec.max(), sec.min()
            if max_value.item() > MAX_FEATURE_VALUE:
                raise Exception(
                    f"A {ftype} feature type has max value {max_value} which is >"
                    f" than accepted post pre-processing max of {MAX_FEATURE_VALUE}")
            elif min_value.item() < MIN_FEATURE_VALUE:
                raise Exception(
                    f"A {ftype} feature type has min value {min_value} which is <"
                    f" accepted post pre-processing min of {MIN_FEATURE_VALUE}")
