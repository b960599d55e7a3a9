"""Scoring + top-k at the shapes of the bench (CUDA events, candidate pool larger than L2): 4096 users x 1 M candidates (one
GPU), 32768 x 125 k (per rank at 8 GPUs), 128 x 1 M (HBM-bound small batch).  Prints TFLOP/s of 2 B N D and GB/s of the
table stream against MEASURED_PEAKS.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unirec_b200 import ops

dev = torch.device("cuda:0")
pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
g = torch.Generator(device="cuda").manual_seed(3)
for B, N in ((4096, 1_000_000), (32768, 125_000), (16384, 250_000), (128, 1_000_000), (256, 1_000_000)):
    u = torch.randn(B, 1024, device=dev, generator=g).to(torch.bfloat16)
    c = torch.randn(N, 1024, device=dev, generator=g).to(torch.bfloat16)
    ci = ops.inv_l2_norm(c)
    for _ in range(3):
        ops.score_topk(u, c, 100, cand_inv=ci)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.score_topk(u, c, 100, cand_inv=ci)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = 2.0 * B * N * 1024 / ms / 1e9
    gbs = N * 1024 * 2 / ms / 1e6
    print(f"score_topk {B:6d} x {N:8d}: {ms:8.3f} ms  {tf:7.1f} TFLOP/s ({tf / pk['bf16_tflops']:.3f} of burst bf16)  "
          f"table stream {gbs:7.1f} GB/s ({gbs / pk['hbm_gbs']:.3f} of HBM)", flush=True)
    del u, c, ci
