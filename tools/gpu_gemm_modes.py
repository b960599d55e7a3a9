"""Epilogue cost of the CTA-pair GEMM under sustained load: the same shape with the bias / bias+erf-GELU / bias+residual
epilogues, each timed over a long back-to-back loop (the part settles at its power-capped clock).  Item-path shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unirec_b200 import ops

dev = torch.device("cuda:0")
bf = torch.bfloat16
M = 131072


def run(N, K, epi, iters=150):
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.03).to(bf)
    b = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).to(bf) if epi == ops.EPI_BIAS_RESIDUAL else None
    out = torch.empty(M, N, device=dev, dtype=bf)
    fn = lambda: ops.linear(a, w, b, epilogue=epi, residual=res, out=out)
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K / ms / 1e9


if __name__ == "__main__":
    names = {ops.EPI_BIAS: "bias", ops.EPI_BIAS_GELU: "bias+gelu", ops.EPI_BIAS_RESIDUAL: "bias+residual"}
    for N, K in ((4096, 1024), (1024, 1024), (1024, 4096), (3072, 1024)):
        for epi in (ops.EPI_BIAS, ops.EPI_BIAS_GELU, ops.EPI_BIAS_RESIDUAL):
            ms, tf = run(N, K, epi)
            print(f"M={M} N={N:5d} K={K:5d} {names[epi]:14s} {ms:7.3f} ms  {tf:7.1f} TFLOP/s", flush=True)
