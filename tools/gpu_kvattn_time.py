"""Time one layer of the fused K/V-projection + attention kernel at the bench shape (512 users x 1600 keys, 16 heads,
E = 1024; CUDA events, 6 launches after warm-up) next to the bare projection GEMM + attention_tc of the materialised path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unirec_b200 import ops

dev = torch.device("cuda:0")
B, S, heads, E = 512, 1600, 16, 1024
H = heads * 64
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(B * S, E, device=dev, generator=g).to(torch.bfloat16)
wk = (torch.randn(H, E, device=dev, generator=g) / 32).to(torch.bfloat16)
wv = (torch.randn(H, E, device=dev, generator=g) / 32).to(torch.bfloat16)
q = torch.randn(B * 64, H, device=dev, generator=g).to(torch.bfloat16)
bv = torch.zeros(H, device=dev)
mask = torch.ones(B, S, device=dev)
wp = ops.pack_kv_weights(wk, wv)
wkv = torch.cat([wk, wv], 0).contiguous()
bkv = torch.zeros(2 * H, device=dev)


def timeit(fn, n=6, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


flops = 2.0 * B * S * 2 * H * E + 4.0 * B * heads * 64 * S * 64
ms = timeit(lambda: ops.kv_attention(x, wp, q, bv, batch=B, num_heads=heads, nk=S, key_mask=mask))
print(f"impl={os.environ.get('UNIREC_KV_ATTENTION_IMPL')} debug={os.environ.get('UNIREC_KV_DEBUG', '0')}: fused {ms:.3f} ms per "
      f"{B} users ({flops / ms / 1e9:.0f} TFLOP/s; x8 = {8 * ms:.1f} ms per 4096 users and layer)", flush=True)
if os.environ.get("UNIREC_KV_DEBUG", "0") == "0":
    def mat():
        kv = ops.linear(x, wkv, bkv)
        return ops.attention(q, kv[:, :H], kv[:, H:], batch=B, num_heads=heads, nq=64, nk=S, key_mask=mask)
    ms2 = timeit(mat)
    ms3 = timeit(lambda: ops.linear(x, wkv, bkv))
    print(f"materialised: projection {ms3:.3f} ms + attention {ms2 - ms3:.3f} ms = {ms2:.3f} ms per {B} users", flush=True)
