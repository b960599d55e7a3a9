"""User cross-attention (attention_tc_kernel, 64 queries x 1600 keys, 512 users = the bench's K/V chunk) timed per launch with
CUDA events (a) alone in a short burst and (b) inside the step's kernel mix - one K/V projection GEMM (819200 x 8192 x 1024)
followed by the four layers' attention launches, repeated until the part sits at its power cap.  One process per variant:
UNIREC_ATTN_EXP_EMU=0..3 (quarters of the softmax exponentials on the FMA pipe).  Also checks the variant against fp32 torch
on 3 users."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unirec_b200 import ops

dev = torch.device("cuda:0")
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
HBM = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
bf = torch.bfloat16
H, heads, Q, S, B, L = 1024, 16, 64, 1600, 512, 4
g = torch.Generator(device=dev).manual_seed(5)
enc = (torch.randn(B * S, H, device=dev, generator=g) * 0.7).to(bf)
w = (torch.randn(2 * H * L, H, device=dev, generator=g) * 0.03).to(bf)
bias = torch.randn(2 * H * L, device=dev, generator=g) * 0.02
qc = (torch.randn(B * Q, H, device=dev, generator=g)).to(bf)
lens = torch.randint(1, 51, (B,), device=dev, generator=g) * 32
lens[0], lens[1] = S, 32
mask = (torch.arange(S, device=dev)[None, :] < lens[:, None]).float().contiguous()
kv = ops.linear(enc, w, bias)
nbytes = (2 * Q + 2 * S) * B * H * 2


def attn(layer, m=mask):
    off = layer * 2 * H
    return ops.attention(qc, kv[:, off:off + H], kv[:, off + H:off + 2 * H], batch=B, num_heads=heads, nq=Q, nk=S, key_mask=m)


# ---- parity on users 0 (full length), 1 (one item), 2 against fp32 torch
out = attn(1).float().view(B, Q, heads, 64)
for u in (0, 1, 2):
    k = kv[u * S:(u + 1) * S, 2 * H:3 * H].float().view(S, heads, 64)
    v = kv[u * S:(u + 1) * S, 3 * H:4 * H].float().view(S, heads, 64)
    q = qc[u * Q:(u + 1) * Q].float().view(Q, heads, 64)
    s = torch.einsum("qhd,khd->hqk", q, k) / 8.0
    s = s.masked_fill(mask[u][None, None, :] == 0, float("-inf"))
    ref = torch.einsum("hqk,khd->qhd", torch.softmax(s, -1), v)
    d = float((out[u] - ref).abs().max())
    print(f"emu={os.environ.get('UNIREC_ATTN_EXP_EMU', 'default')} user {u}: max|d| vs fp32 = {d:.5f} (|ref|max {float(ref.abs().max()):.3f})")
    assert d < 0.03


def events(n):
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


for name, m in (("ragged mask", mask), ("no mask", None)):
    for _ in range(3):
        attn(0, m)
    ev = events(12)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(ev):
        a.record(); attn(i % L, m); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
    print(f"alone, {name:12s}: {ms:.3f} ms  {nbytes / ms / 1e6:7.1f} GB/s  {100 * nbytes / ms / 1e6 / HBM:5.1f}% of HBM peak", flush=True)

# ---- inside the step's mix, sustained (~2.5 s)
reps = 48
ev = events(reps * L)
gev = events(reps)
torch.cuda.synchronize()
for r in range(reps):
    gev[r][0].record(); ops.linear(enc, w, bias, out=kv) if False else ops.linear(enc, w, bias); gev[r][1].record()
    for l in range(L):
        a, b = ev[r * L + l]
        a.record(); attn(l); b.record()
torch.cuda.synchronize()
tail = [a.elapsed_time(b) for a, b in ev[len(ev) // 2:]]
gt = [a.elapsed_time(b) for a, b in gev[reps // 2:]]
ms = sum(tail) / len(tail)
print(f"in the mix (second half of {reps} GEMM + 4 attention rounds): attention {ms:.3f} ms  {nbytes / ms / 1e6:7.1f} GB/s  "
      f"{100 * nbytes / ms / 1e6 / HBM:5.1f}% of HBM peak; GEMM {sum(gt) / len(gt):.3f} ms", flush=True)
