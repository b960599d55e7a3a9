"""score_topk (cosine top-100 over 1 M candidates) timed on differently distributed candidate tables: i.i.d. rows, rows that
repeat with period 16384 (what bench.py's item generation produced while it cycled over four 4096-item field batches: every
distinct item ~61 times in the pool, i.e. 61-fold ties at every rank), and i.i.d. rows around a common mean."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unirec_b200 import ops

dev = torch.device("cuda:0")
bf = torch.bfloat16
N, D, k = 1_000_000, 1024, 100
g = torch.Generator(device=dev).manual_seed(3)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


base = torch.randn(N, D, device=dev, generator=g)
tables = {
    "i.i.d. rows": base.to(bf),
    "rows repeat with period 16384": base[torch.arange(N, device=dev) % 16384].to(bf),
    "i.i.d. rows + common mean (|mean| = 3 sigma)": (base * 0.3 + torch.randn(D, device=dev, generator=g)).to(bf),
}
del base
for name, c in tables.items():
    ci = ops.inv_l2_norm(c)
    for B in (128, 4096):
        u = torch.randn(B, D, device=dev, generator=g).to(bf)
        ms = timeit(lambda: ops.score_topk(u, c, k, cand_inv=ci))
        print(f"{name:48s} B={B:5d}: {ms:8.3f} ms  {2 * B * N * D / ms / 1e9:8.1f} TFLOP/s  table stream {N * D * 2 / ms / 1e6:7.1f} GB/s",
              flush=True)
