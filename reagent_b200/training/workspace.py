"""Device workspaces for the fused trainers (activations, dZ, gradient partials)."""
import torch

from .. import _lib
from ..models.arena import ParamArena


class NetWorkspace:
    """hidden[l] / dz[l] / input buffers of one network for a fixed batch size."""

    def __init__(self, arena: ParamArena, batch: int, device, need_input: bool = False):
        L = len(arena.acts)
        self.arena = arena
        self.batch = batch
        self.hidden = [torch.empty(batch, arena.dims[l + 1], device=device) for l in range(L - 1)]
        self.dz = [torch.empty(batch, arena.dims[l + 1], device=device) for l in range(L)]
        self.input = torch.empty(batch, arena.dims[0], device=device) if need_input else None
        self.c = _lib.NetWsT()
        for l, t in enumerate(self.hidden):
            self.c.hidden[l] = t.data_ptr()
        for l, t in enumerate(self.dz):
            self.c.dz[l] = t.data_ptr()
        self.c.input = None if self.input is None else self.input.data_ptr()


def ensure_gpart(arena: ParamArena, splits: int):
    """[splits, n] gradient partial slab (zeroed once: alignment padding is never written)."""
    flat = arena.flat
    if arena.gpart is None or arena.gpart.shape[0] != splits or arena.gpart.device != flat.device:
        arena.gpart = torch.zeros(splits, arena.n, device=flat.device)
    return arena.gpart


def wgrad(arena: ParamArena, ws: NetWorkspace, net_input, batch: int):
    """Launch the split-K weight-gradient kernel for one network."""
    splits = _lib.lib().rb200_wgrad_splits_for(arena.desc(), batch)
    g = ensure_gpart(arena, splits)
    rc = _lib.lib().rb200_mlp_wgrad(arena.desc(), _lib.ptr(net_input), batch, ws.c,
                                    g.data_ptr(), splits, _lib.cur_stream())
    _lib.check(rc, "rb200_mlp_wgrad")
    arena.grad_ready = True


def head_backward_dx(arena: ParamArena, ws: NetWorkspace, batch: int, cache: dict):
    """dZ of the layer below a wide head: dz[L-2] = (dz[L-1] . W_head) * act'(h[L-2])
    (torch.nn.functional.linear's backward w.r.t. its input).  Wide heads (QR-DQN: A*N atoms,
    C51: A*51) take the tcgen05 split-K path, which needs a scratch buffer kept in `cache`."""
    lib, st = _lib.lib(), _lib.cur_stream()
    L = len(arena.acts)
    K, N = arena.dims[L - 1], arena.dims[L]
    W = arena.flat.data_ptr() + 4 * arena.w_off[L - 1]
    dz, h, out = ws.dz[L - 1], ws.hidden[L - 2], ws.dz[L - 2]
    key = ("dx_scratch", K, N, batch)
    if key not in cache:
        nbytes = int(lib.rb200_linear_backward_dx_tc_scratch_bytes(K, N, batch))
        cache[key] = torch.empty(nbytes // 4, device=arena.flat.device) if nbytes else None
    scratch = cache[key]
    if scratch is not None:
        rc = lib.rb200_linear_backward_dx_tc(W, K, N, dz.data_ptr(), h.data_ptr(), arena.acts[L - 2],
                                             batch, out.data_ptr(), scratch.data_ptr(),
                                             scratch.numel() * 4, st)
        _lib.check(rc, "rb200_linear_backward_dx_tc")
    else:
        rc = lib.rb200_linear_backward_dx(W, K, N, dz.data_ptr(), h.data_ptr(), arena.acts[L - 2],
                                          batch, out.data_ptr(), st)
        _lib.check(rc, "rb200_linear_backward_dx")


def reduced_grad(arena: ParamArena) -> torch.Tensor:
    """Flat gradient = fixed-order sum of the partials (for inspection / all-reduce)."""
    assert arena.gpart is not None, "no gradient partials computed yet"
    out = torch.empty(arena.n, device=arena.flat.device)
    rc = _lib.lib().rb200_grad_reduce(arena.gpart.data_ptr(), arena.gpart.shape[0], arena.n,
                                      out.data_ptr(), _lib.cur_stream())
    _lib.check(rc, "rb200_grad_reduce")
    return out


def param_grads(arena: ParamArena, params):
    """Per-parameter gradient views (same shapes as the parameters)."""
    g = reduced_grad(arena)
    base = arena.flat.data_ptr()
    out = []
    for p in params:
        off = (p.data_ptr() - base) // 4
        out.append(g[off:off + p.numel()].view_as(p))
    return out
