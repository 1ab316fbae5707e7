import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reagent_b200 import _lib
B, K, N = 4096, 128, 6400
x = torch.randn(B, K, device="cuda"); W = torch.randn(N, K, device="cuda") / 11; b = torch.zeros(N, device="cuda")
out = torch.empty(B, N, device="cuda")
for _ in range(4):
    _lib.check(_lib.lib().rb200_linear_forward(W.data_ptr(), b.data_ptr(), 0, K, N, x.data_ptr(), B, out.data_ptr(), _lib.cur_stream()))
torch.cuda.synchronize(); print("ok")
