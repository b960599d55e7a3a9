"""Per-kernel timing of the HBM-bound kernels at the BASELINE shapes (CUDA events, inputs larger than L2 or
rotated over several buffers).  Prints achieved GB/s of ALGORITHMIC bytes against MEASURED_PEAKS.json."""
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


def timeit(fn, n=10, warm=3):
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


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:52s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / HBM:5.1f}% of measured HBM peak", flush=True)


def attention_cases():
    H, heads = 1024, 16
    # item self-attention: B x 32 queries, fused qkv buffer [B*32, 3072]
    B = 8192
    qkv = torch.randn(B * 32, 3 * H, device=dev).to(bf)
    ms = timeit(lambda: ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], batch=B, num_heads=heads, nq=32, nk=32))
    report("attention item self  (B=8192, 32q x 32k)", ms, 4 * B * 32 * H * 2)
    # item cross: 32 queries x 14 keys, kv_all [B*14, 12*1024]
    qc = torch.randn(B * 32, H, device=dev).to(bf)
    kv = torch.randn(B * 14, 12 * H, device=dev).to(bf)
    mask = torch.ones(B, 14, device=dev)
    ms = timeit(lambda: ops.attention(qc, kv[:, :H], kv[:, H:2 * H], batch=B, num_heads=heads, nq=32, nk=14, key_mask=mask))
    report("attention item cross (B=8192, 32q x 14k)", ms, (2 * 32 + 2 * 14) * B * H * 2)
    # user self: 64 x 64
    Bu = 2048
    qkv = torch.randn(Bu * 64, 3 * H, device=dev).to(bf)
    ms = timeit(lambda: ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], batch=Bu, num_heads=heads, nq=64, nk=64))
    report("attention user self  (B=2048, 64q x 64k)", ms, 4 * Bu * 64 * H * 2)
    # user cross: 64 x 1600, kv_all [B*1600, 8*1024]
    Bu = 256
    qc = torch.randn(Bu * 64, H, device=dev).to(bf)
    kv = torch.randn(Bu * 1600, 8 * H, device=dev).to(bf)
    mask = torch.ones(Bu, 1600, device=dev)
    ms = timeit(lambda: ops.attention(qc, kv[:, 2 * H:3 * H], kv[:, 3 * H:4 * H], batch=Bu, num_heads=heads, nq=64, nk=1600,
                                      key_mask=mask))
    report("attention user cross (B=256, 64q x 1600k)", ms, (2 * 64 + 2 * 1600) * Bu * H * 2)


def rowwise_cases():
    H = 1024
    M = 131072
    x = torch.randn(M, H, device=dev).to(bf)
    g = torch.ones(H, device=dev)
    b = torch.zeros(H, device=dev)
    out = torch.empty_like(x)
    ms = timeit(lambda: ops.layernorm(x, g, b, 1e-12, out=out))
    report("layernorm bf16->bf16 (131072 x 1024)", ms, 2 * M * H * 2)
    xf = torch.randn(M, H, device=dev)
    ms = timeit(lambda: ops.layernorm(xf, g, b, 1e-12, out=out))
    report("layernorm fp32->bf16 (131072 x 1024)", ms, M * H * 6)
    f = torch.randn(4096 * 14, 1024, device=dev)
    ms = timeit(lambda: ops.cast_bf16(f))
    report("cast fp32->bf16 (57344 x 1024)", ms, f.numel() * 6)
    tok = torch.randn(8192, 32, H, device=dev).to(bf)
    ms = timeit(lambda: ops.mean_tokens(tok))
    report("mean_tokens (8192 x 32 x 1024)", ms, tok.numel() * 2 + 8192 * H * 2)
    table = torch.randn(200000, 32, H, device=dev).to(bf)
    hist = torch.randint(0, 200000, (256, 50), device=dev)
    lens = torch.full((256,), 50, device=dev, dtype=torch.int32)
    ms = timeit(lambda: ops.build_user_sequence(table, hist, lens))
    report("build_user_sequence (256 users x 50 x 32 x 1024)", ms, 2 * 256 * 1600 * H * 2)
    c = torch.randn(1_000_000, H, device=dev).to(bf)
    ms = timeit(lambda: ops.inv_l2_norm(c))
    report("inv_l2_norm (1M x 1024)", ms, c.numel() * 2)


def scoring_cases():
    D = 1024
    c = torch.randn(1_000_000, D, device=dev).to(bf)
    ci = ops.inv_l2_norm(c)
    for B in (128, 1024, 4096):
        u = torch.randn(B, D, device=dev).to(bf)
        ms = timeit(lambda: ops.score_topk(u, c, 100, cand_inv=ci), n=5)
        print(f"score_topk B={B:5d} N=1M: {ms:8.3f} ms  {2 * B * 1e6 * D / ms / 1e9:8.1f} TFLOP/s  "
              f"table stream {c.numel() * 2 / ms / 1e6:8.1f} GB/s ({100 * c.numel() * 2 / ms / 1e6 / HBM:5.1f}% HBM)", flush=True)


def train_cases():
    """Kernels of the cfg-2 training step at its shapes (1024 items x 32 queries, hidden 1024, 16 heads)."""
    H, heads, B, Q, F = 1024, 16, 1024, 32, 14
    M = B * Q
    drop = (ops.dropout_threshold(0.2), 12345, 7)
    x = torch.randn(M, H, device=dev).to(bf)
    res = torch.randn(M, H, device=dev).to(bf)
    ms = timeit(lambda: ops.dropout_add(x, res, drop))
    report("dropout_add (32768 x 1024)", ms, 3 * M * H * 2)
    ms = timeit(lambda: ops.dropout_backward(x, drop))
    report("dropout_backward (32768 x 1024)", ms, 2 * M * H * 2)
    qkv = torch.randn(M, 3 * H, device=dev).to(bf)
    do = torch.randn(M, H, device=dev).to(bf)
    dqkv = torch.empty_like(qkv)
    for name, d in (("", None), (" +dropout", drop)):
        ms = timeit(lambda: ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], batch=B, num_heads=heads, nq=Q, nk=Q,
                                          dropout=d))
        report(f"attention fwd self 32x32{name} (B=1024)", ms, 4 * M * H * 2)
        ms = timeit(lambda: ops.attention_backward(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], do, dqkv[:, :H],
                                                   dqkv[:, H:2 * H], dqkv[:, 2 * H:], batch=B, num_heads=heads, nq=Q, nk=Q,
                                                   dropout=d))
        report(f"attention bwd self 32x32{name} (B=1024)", ms, 8 * M * H * 2)
    kv = torch.randn(B * F, 12 * H, device=dev).to(bf)
    dkv = torch.zeros_like(kv)
    qc = torch.randn(M, H, device=dev).to(bf)
    dqc = torch.empty_like(qc)
    mask = torch.ones(B, F, device=dev)
    for name, d in (("", None), (" +dropout", drop)):
        ms = timeit(lambda: ops.attention(qc, kv[:, :H], kv[:, H:2 * H], batch=B, num_heads=heads, nq=Q, nk=F, key_mask=mask,
                                          dropout=d))
        report(f"attention fwd cross 32x14{name} (B=1024)", ms, (2 * Q + 2 * F) * B * H * 2)
        ms = timeit(lambda: ops.attention_backward(qc, kv[:, :H], kv[:, H:2 * H], do, dqc, dkv[:, :H], dkv[:, H:2 * H],
                                                   batch=B, num_heads=heads, nq=Q, nk=F, key_mask=mask, dropout=d))
        report(f"attention bwd cross 32x14{name} (B=1024)", ms, (4 * Q + 4 * F) * B * H * 2)
    g = torch.ones(H, device=dev)
    dg, db = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    ms = timeit(lambda: ops.layernorm_backward(x, do, g, 1e-12, dg, db, dy2=res))
    report("layernorm_backward (32768 x 1024, dy + dy2)", ms, 4 * M * H * 2)
    z = torch.randn(M, 4 * H, device=dev).to(bf)
    ms = timeit(lambda: ops.gelu(z))
    report("gelu fwd (32768 x 4096)", ms, 2 * M * 4 * H * 2)
    ms = timeit(lambda: ops.gelu_backward(z, z))
    report("gelu bwd (32768 x 4096)", ms, 3 * M * 4 * H * 2)
    bsum = torch.zeros(4 * H, device=dev)
    ms = timeit(lambda: ops.colsum(z, bsum))
    report("colsum (32768 x 4096)", ms, M * 4 * H * 2)
    w = torch.randn(4 * H, H, device=dev).to(bf)
    dw = torch.zeros(4 * H, H, device=dev)
    for name, fn, fl in (("dgrad 32768x1024x4096", lambda: ops.linear_dgrad(z, w), 2.0 * M * H * 4 * H),
                         ("wgrad 4096x1024x32768", lambda: ops.linear_wgrad(z, x, dw), 2.0 * M * H * 4 * H),
                         ("fwd   32768x4096x1024", lambda: ops.linear(x, w), 2.0 * M * H * 4 * H)):
        ms = timeit(fn)
        print(f"gemm {name:44s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


def joint_cases():
    """Per-user candidate-list scoring (csrc/list_scoring.cu): one streaming pass over [B, 1 + C, D] rows."""
    from unirec_b200.joint import InfoNCELoss
    for dt, name in ((torch.float32, "fp32"), (bf, "bf16")):
        B, C, D = 8192, 100, 1024          # 3.4 GB fp32 / 1.7 GB bf16 of candidate rows: larger than L2
        u = torch.randn(B, D, device=dev).to(dt)
        p = torch.randn(B, D, device=dev).to(dt)
        n = torch.randn(B, C, D, device=dev).to(dt)
        m = torch.rand(B, C, device=dev) < 0.9
        es = u.element_size()
        ms = timeit(lambda: ops.list_scores(u, p, n, mask=m))
        report(f"list_scores {name} (B=8192, 1+100 x 1024, 90% valid)", ms, (B * 2 + int(m.sum())) * D * es)
        ms = timeit(lambda: ops.list_scores(u, p, n))
        report(f"list_scores {name} (B=8192, 1+100 x 1024, no mask)", ms, B * (C + 2) * D * es)
        sims, inv = ops.list_scores(u, p, n)
        ms = timeit(lambda: ops.infonce_rank(sims, 0.07))
        report(f"infonce_rank (B=8192, 101 sims)", ms, B * (C + 1) * 4)
        dl = torch.full((B,), 1.0 / B, device=dev)
        ms = timeit(lambda: ops.list_scores_backward(u, p, n, sims, inv, dl, 0.07))
        report(f"list_scores_backward {name} (d_user only)", ms, B * (C + 2) * D * es + B * D * 4)
        crit = InfoNCELoss(0.07)
        ms = timeit(lambda: crit(u, p, n, m))
        report(f"InfoNCELoss.forward {name} (module call, masked)", ms, (B * 2 + int(m.sum())) * D * es)
    B, S, nh, Q, Hd = 64, 2048, 10, 32, 1024
    ids = torch.randint(0, 150_000, (B, S), device=dev)
    tok_ids = 151_700 + torch.arange(nh * Q, device=dev)
    ids[:, 100:100 + nh * Q] = tok_ids
    text = torch.randn(B, S, Hd, device=dev).to(bf)
    toks = torch.randn(B, nh * Q, Hd, device=dev)
    ms = timeit(lambda: ops.inject_tokens(text, ids, tok_ids, toks))
    report("inject_tokens (B=64, S=2048, 320 placeholders, fp32->bf16)", ms, B * S * 8 + B * nh * Q * Hd * 6)


if __name__ == "__main__":
    what = sys.argv[1:] or ["attention", "rowwise", "scoring", "joint"]
    print(torch.cuda.get_device_name(0), "HBM peak", HBM, "GB/s")
    if "attention" in what:
        attention_cases()
    if "rowwise" in what:
        rowwise_cases()
    if "scoring" in what:
        scoring_cases()
    if "joint" in what:
        joint_cases()
    if "train" in what:
        train_cases()
