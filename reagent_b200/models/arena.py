"""Flat parameter arena shared by every model in this package.

Each network keeps ordinary `nn.Linear` sub-modules (so `state_dict()` keys and shapes are
the reference's: `fc.dnn.{i}.0.weight`, reagent/models/fully_connected_network.py:101-153)
but their `.data` are views into ONE contiguous fp32 buffer laid out
[W0, b0, W1, b1, ...] with every tensor starting on a 16-byte boundary.  The CUDA kernels
see the arena (base pointer + offsets, `rb200_mlp_t`); Adam, the Polyak update and the
gradient all-reduce are single launches over it.
"""
import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _lib


def _align4(n: int) -> int:
    return (n + 3) & ~3


class ParamArena:
    """The flat buffer + layout of one network (or of one stand-alone parameter)."""

    def __init__(self, dims: List[int], acts: List[int]):
        assert len(dims) == len(acts) + 1
        assert 1 <= len(acts) <= _lib.MAX_LAYERS, f"at most {_lib.MAX_LAYERS} layers are supported"
        self.dims = list(dims)
        self.acts = list(acts)
        self.w_off, self.b_off = [], []
        off = 0
        for i in range(len(acts)):
            self.w_off.append(off)
            off = _align4(off + dims[i] * dims[i + 1])
            self.b_off.append(off)
            off = _align4(off + dims[i + 1])
        self.n = off
        self.flat: Optional[torch.Tensor] = None
        # filled by the trainer that owns the update of this arena
        self.gpart: Optional[torch.Tensor] = None   # [splits, n] gradient partials
        self.grad_ready = False
        self._desc = None

    # -- hooks for arenas with derived regions (models/dueling_q_network.py) ------
    def refresh(self):
        """Bring derived parameters up to date before the kernels read the arena (no-op here)."""

    def finish_grads(self):
        """Map gradients of derived parameters back onto the true ones (no-op here)."""

    # -- views ---------------------------------------------------------------
    def weight_view(self, flat, l):
        o, i = self.dims[l + 1], self.dims[l]
        return flat[self.w_off[l]: self.w_off[l] + o * i].view(o, i)

    def bias_view(self, flat, l):
        o = self.dims[l + 1]
        return flat[self.b_off[l]: self.b_off[l] + o]

    def desc(self, n_layers=None) -> _lib.MlpT:
        """ctypes descriptor for the current flat buffer (optionally only the first
        `n_layers` layers, e.g. the trunk below a wide head)."""
        assert self.flat is not None
        L = len(self.acts) if n_layers is None else n_layers
        d = _lib.MlpT()
        d.n_layers = L
        for i, v in enumerate(self.dims[: L + 1]):
            d.dims[i] = v
        for i, a in enumerate(self.acts[:L]):
            d.act[i] = a
            d.w_off[i] = self.w_off[i]
            d.b_off[i] = self.b_off[i]
        d.params = self.flat.data_ptr()
        d.n_params = self.n
        return d


def flatten_linears(linears: List[nn.Linear], arena: ParamArena, device=None) -> torch.Tensor:
    """Move the Linear parameters into one flat buffer (keeping their values) and re-point
    `.data` at views of it.  Returns the flat tensor."""
    dev = device if device is not None else linears[0].weight.device
    flat = torch.zeros(arena.n, dtype=torch.float32, device=dev)
    for l, lin in enumerate(linears):
        w = arena.weight_view(flat, l)
        b = arena.bias_view(flat, l)
        w.copy_(lin.weight.data.to(dev, torch.float32))
        b.copy_(lin.bias.data.to(dev, torch.float32))
        lin.weight.data = w
        lin.bias.data = b
        lin.weight._rb200_arena = arena
        lin.bias._rb200_arena = arena
    arena.flat = flat
    arena.gpart = None
    arena.grad_ready = False
    return flat


def arena_of(params) -> ParamArena:
    """The arena a list of parameters belongs to (all must share one)."""
    params = list(params)
    if not params:
        raise ValueError("empty parameter list")
    a = getattr(params[0], "_rb200_arena", None)
    if a is None:
        raise ValueError(
            "parameter does not belong to a reagent_b200 network (no flat arena); "
            "build the network with reagent_b200.models / net_builder")
    for p in params:
        if getattr(p, "_rb200_arena", None) is not a:
            raise ValueError("parameters of one optimizer must come from ONE reagent_b200 network")
    return a


class ScalarArena(ParamArena):
    """Arena wrapping one stand-alone contiguous parameter (e.g. SAC's log_alpha).  `flat`
    follows the parameter when the owning module is moved between devices."""

    def __init__(self, param: torch.nn.Parameter):
        self.dims, self.acts, self.w_off, self.b_off = [], [], [], []
        self.n = param.numel()
        self._param = param
        self.gpart = None
        self.grad_ready = False
        param._rb200_arena = self

    @property
    def flat(self):
        return self._param.data.view(-1)

    @flat.setter
    def flat(self, v):
        pass
