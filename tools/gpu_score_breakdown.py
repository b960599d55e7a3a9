"""One score_topk call per shape for an `ncu --metrics gpu__time_duration.sum` launch list (per-kernel durations)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unirec_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(3)
for B, N in ((4096, 1_000_000), (32768, 125_000)):
    u = torch.randn(B, 1024, device=dev, generator=g).to(torch.bfloat16)
    c = torch.randn(N, 1024, device=dev, generator=g).to(torch.bfloat16)
    ci = ops.inv_l2_norm(c)
    ops.score_topk(u, c, 100, cand_inv=ci)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ops.score_topk(u, c, 100, cand_inv=ci)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    del u, c, ci
