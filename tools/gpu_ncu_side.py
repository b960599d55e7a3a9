"""One launch of each side kernel (list scoring, token injection, LayerNorm backward, evaluation metrics) at the shapes
of tools/gpu_bench_kernels.py, for an `ncu -k regex:...` capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unirec_b200 import ops

dev = torch.device("cuda:0")
B, C, D = 8192, 100, 1024
u = torch.randn(B, D, device=dev)
p = torch.randn(B, D, device=dev)
n = torch.randn(B, C, D, device=dev)
m = torch.rand(B, C, device=dev) < 0.9
sims, inv = ops.list_scores(u, p, n, mask=m)
ops.infonce_rank(sims, 0.07)
ops.list_scores_backward(u, p, n, sims, inv, torch.full((B,), 1.0 / B, device=dev), 0.07, mask=m)
ub, pb, nb = u.bfloat16(), p.bfloat16(), n.bfloat16()
ops.list_scores(ub, pb, nb, mask=m)
ids = torch.randint(0, 150_000, (64, 2048), device=dev)
tok_ids = 151_700 + torch.arange(320, device=dev)
ids[:, 100:420] = tok_ids
ops.inject_tokens(torch.randn(64, 2048, 1024, device=dev).bfloat16(), ids, tok_ids, torch.randn(64, 320, 1024, device=dev))
M, H = 32768, 1024
x = torch.randn(M, H, device=dev).bfloat16()
dy = torch.randn(M, H, device=dev).bfloat16()
g = torch.ones(H, device=dev)
ops.layernorm_backward(x, dy, g, 1e-12, torch.zeros(H, device=dev), torch.zeros(H, device=dev), dy2=x,
                       dropout=(ops.dropout_threshold(0.2), 1, 2), dbias=torch.zeros(H, device=dev))
rec = torch.randn(4096, 14, 1024, device=dev)
ops.reconstruction_metrics(rec, torch.randn(4096, 14, 1024, device=dev), torch.ones(4096, 14, device=dev))
torch.cuda.synchronize()
print("ok")
