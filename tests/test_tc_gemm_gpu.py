"""tcgen05 / TMEM wide Linear forward (rb200_linear_forward -> tc_linear_fwd_kernel) vs an
fp64 torch reference: full tiles, ragged rows / columns / K, scalar-load path, activations."""
import pytest
import torch

from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _run(B, K, N, act):
    from reagent_b200 import _lib

    g = torch.Generator().manual_seed(B * 7 + K * 3 + N)
    x = torch.randn(B, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    ref = x.double() @ W.double().t() + b.double()
    if act == "relu":
        ref = torch.relu(ref)
    elif act == "tanh":
        ref = torch.tanh(ref)
    xd, Wd, bd = x.cuda(), W.cuda().contiguous(), b.cuda()
    out = torch.empty(B, N, device="cuda")
    rc = _lib.lib().rb200_linear_forward(Wd.data_ptr(), bd.data_ptr(), _lib.ACT[act if act else "linear"],
                                         K, N, xd.data_ptr(), B, out.data_ptr(), _lib.cur_stream())
    _lib.check(rc, "rb200_linear_forward")
    torch.cuda.synchronize()
    return G.rel_err(out, ref.float())


@pytest.mark.parametrize("B,K,N,act", [
    (128, 32, 128, None),        # exactly one tile, one k-chunk
    (128, 128, 256, "relu"),     # 4 chunks (ring wraps), 2 column tiles
    (4096, 128, 6400, None),     # QR-DQN head of BASELINE config 3
    (300, 100, 200, "tanh"),     # ragged rows / cols, K not a multiple of 32
    (256, 36, 130, None),        # K tail quad, 2 columns in the last tile
    (129, 7, 129, "relu"),       # scalar-load path (K % 4 != 0)
    (64, 128, 6400, None),       # batch < 128 -> row-tile mma.sync path
])
def test_linear_forward(B, K, N, act):
    assert _run(B, K, N, act) < 1e-5
