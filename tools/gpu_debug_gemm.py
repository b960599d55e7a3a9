"""Diagnostic run of the tcgen05 GEMM on a GPU box: prints error structure, not just pass/fail."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unirec_b200 import ops

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)


def run(M, N, K, block_n, epi=0):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    out = ops.linear(a, w, b, block_n=block_n, out_dtype=torch.float32, epilogue=epi)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + b
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    err = (out - ref).abs()
    bad = err > 5e-3
    print(f"M={M} N={N} K={K} bn={block_n} epi={epi}: max_err={float(err.max()):.4e} bad_frac={float(bad.float().mean()):.4f}",
          flush=True)
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print("  bad rows:", rows[:16].tolist(), "... n=", len(rows), " bad cols:", cols[:16].tolist(), "... n=", len(cols))
        print("  bad by row%8:", [int(bad[i::8].sum()) for i in range(8)])
        print("  bad by col%64 (first 16):", [int(bad[:, i::64].sum()) for i in range(16)])
        r, c = int(rows[0]), int(cols[0])
        print("  sample out/ref:", out[r, c:c + 6].tolist(), ref[r, c:c + 6].tolist())
        # hypothesis: bias only (no accumulation)
        print("  max|out - bias|:", float((out - b).abs().max()))
    return float(err.max())


for shape in [(128, 128, 64, 128), (128, 256, 64, 256), (128, 128, 128, 128), (128, 128, 512, 128),
              (256, 512, 1024, 256), (512, 1024, 1024, 128), (4096, 4096, 1024, 256), (777, 1024, 4096, 256)]:
    run(*shape)
run(512, 4096, 1024, 256, epi=1)

# quick timing
M, N, K = 131072, 1024, 1024
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(N, device=dev)
for bn in (128, 256, 2):
    for (n, k, epi) in [(1024, 1024, 0), (4096, 1024, 1), (3072, 1024, 0), (1024, 1024, 2), (1024, 4096, 2), (8192, 1024, 0)]:
        w = (torch.randn(n, k, device=dev) * 0.05).to(torch.bfloat16)
        b = torch.randn(n, device=dev)
        out = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
        a = torch.randn(M, k, device=dev).to(torch.bfloat16)
        res = torch.randn(M, n, device=dev).to(torch.bfloat16) if epi == 2 else None
        for _ in range(3):
            ops.linear(a, w, b, block_n=bn, out=out, epilogue=epi, residual=res)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.linear(a, w, b, block_n=bn, out=out, epilogue=epi, residual=res)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"timing M={M} N={n} K={k} bn={bn} epi={epi}: {ms:.3f} ms  {2*M*n*k/ms/1e9:.1f} TFLOP/s", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a = torch.randn(M, 1024, device=dev).to(torch.bfloat16)
w = (torch.randn(4096, 1024, device=dev) * 0.05).to(torch.bfloat16)
for _ in range(3): torch.matmul(a, w.t())
e0.record()
for _ in range(10): torch.matmul(a, w.t())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"cuBLAS M={M} N=4096 K=1024: {ms:.3f} ms {2*M*4096*1024/ms/1e9:.1f} TFLOP/s")
