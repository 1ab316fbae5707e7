"""tc_linear_fwd_kernel (through rb200_linear_forward): time vs K and N at B=4096 -> fixed cost per
CTA vs cost per 32-k chunk.  Usage: python profiles/time_tc_gemm.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reagent_b200 import _lib

lib, dev = _lib.lib(), torch.device("cuda", 0)


def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B = 4096
for K, N in ((32, 6400), (64, 6400), (128, 6400), (256, 6400), (512, 6400), (128, 1600), (128, 3200), (128, 12800)):
    x = torch.randn(B, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    out = torch.empty(B, N, device=dev)
    st = _lib.cur_stream()
    us = timeit(lambda: lib.rb200_linear_forward(W.data_ptr(), b.data_ptr(), 0, K, N, x.data_ptr(), B, out.data_ptr(), st))
    ctas = (B // 128) * ((N + 127) // 128)
    print(json.dumps({"K": K, "N": N, "us": round(us, 1), "ctas": ctas, "chunks_per_cta": K // 32,
                      "tflops_alg": round(2 * B * K * N / us / 1e6, 1),
                      "out_GBps": round(B * N * 4 / us / 1e3, 0)}), flush=True)
