"""One small launch of every kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck`
("compute-sanitizer on each kernel's unit test").  Shapes are tiny (a sanitizer slows a kernel
10-100x) but chosen to hit the interesting paths: ragged last tiles, several tiles per CTA (ring wrap-around, both TMEM
accumulator stages), masked / all-masked attention rows, user boundaries inside a K/V tile, list compaction in the
scoring epilogue.  Results are also compared with torch so that a sanitizer run doubles as a smoke test.

Usage (tools/gpu_sanitizer.sh):  compute-sanitizer --tool memcheck python tools/sanitizer_cases.py [case ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unirec_b200 import ops

dev = torch.device("cuda:0")
bf = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(5)


def rnd(*shape, scale=1.0, dtype=bf):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


def close(a, b, tol, what):
    """|a - b| <= tol + 2^-7 |b|: an absolute part for the accumulation and one bf16 output rounding."""
    err = (a.float() - b.float()).abs()
    print(f"  {what}: max|d| = {float(err.max()):.4g}", flush=True)
    assert bool((err <= tol + 2.0 ** -7 * b.float().abs()).all()), (what, float(err.max()))


def case_gemm_cg2():
    # CTA-pair kernel: 3 x 2 tiles on <= 2 clusters (ring wrap-around, both accumulator stages), ragged M
    M, N, K = 600, 512, 256
    a, w, b = rnd(M, K), rnd(N, K, scale=0.06), rnd(N, dtype=torch.float32)
    ref = a.float() @ w.float().t() + b
    close(ops.linear(a, w, b, max_ctas=4), ref, 0.05, "cg2 bias")
    res = rnd(M, N)
    close(ops.linear(a, w, b, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, max_ctas=4), ref + res.float(), 0.06, "cg2 residual")
    close(ops.linear(a, w, b, epilogue=ops.EPI_BIAS_GELU), torch.nn.functional.gelu(ref), 0.05, "cg2 gelu")


def case_gemm_ln():
    # LayerNorm folded into the GEMMs (unirec_linear_ln_bf16): producer statistics, both consumer forms, ragged M
    M, H, I = 300, 256, 512
    eps = 1e-12
    x, w2, b2 = rnd(M, I), rnd(H, I, scale=0.04), rnd(H, dtype=torch.float32)
    res = rnd(M, H)
    gamma = 1.0 + rnd(H, scale=0.1, dtype=torch.float32)
    beta = rnd(H, scale=0.1, dtype=torch.float32)
    st = ops.ln_stats_buffer(M, H, dev)
    pre = ops.linear_ln(x, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, stats_out=st, eps=eps, hidden=H)
    close(pre, x.float() @ w2.float().t() + b2 + res.float(), 0.06, "ln producer")
    h = torch.nn.functional.layer_norm(pre.float(), (H,), gamma, beta, eps)
    w1, b1 = rnd(I, H, scale=0.04, dtype=torch.float32), rnd(I, dtype=torch.float32)
    wf, bf_, cf = ops.fold_layernorm_weights(w1, b1, gamma, beta)
    y = ops.linear_ln(pre, wf, bf_, epilogue=ops.EPI_BIAS_GELU, ln_in=(st, cf), eps=eps, hidden=H)
    close(y, torch.nn.functional.gelu(h @ w1.t() + b1), 0.06, "ln consumer (A operand, gelu)")
    st2 = ops.ln_stats_buffer(M, H, dev)
    pre2 = ops.linear_ln(x, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, residual=pre, ln_res=(st, gamma, beta), stats_out=st2,
                         eps=eps, hidden=H)
    close(pre2, x.float() @ w2.float().t() + b2 + h, 0.06, "ln consumer (residual)")


def case_gemm_1cta():
    M, N, K = 200, 384, 192
    a, w, b = rnd(M, K), rnd(N, K, scale=0.06), rnd(N, dtype=torch.float32)
    ref = a.float() @ w.float().t() + b
    close(ops.linear(a, w, b, out_dtype=torch.float32, block_n=128), ref, 2e-3 + 1e-4 * float(ref.abs().max()), "1cta fp32 out")
    close(ops.linear(a, w, b, block_n=128, max_ctas=2), ref, 0.05, "1cta bf16 out")


def case_gemm_grad():
    rows, N, K = 264, 256, 128
    dy, x, w = rnd(rows, N, scale=0.1), rnd(rows, K), rnd(N, K, scale=0.06)
    close(ops.linear_dgrad(dy, w), dy.float() @ w.float(), 0.05, "dgrad")
    dw = torch.zeros(N, K, device=dev)
    ops.linear_wgrad(dy, x, dw)
    close(dw, dy.float().t() @ x.float(), 0.05, "wgrad")


def _attn_ref(q, k, v, mask, B, heads, nq, nk, q_b=False):
    qf = (q.float().view(1 if q_b else B, nq, heads, 64).permute(0, 2, 1, 3)).expand(B, -1, -1, -1)
    kf = k.float().view(B, nk, heads, 64).permute(0, 2, 1, 3)
    vf = v.float().view(B, nk, heads, 64).permute(0, 2, 1, 3)
    s = qf @ kf.transpose(-1, -2) / 8.0
    if mask is not None:
        s = s + (1.0 - mask)[:, None, None, :] * torch.finfo(torch.float32).min
    return (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(B * nq, heads * 64)


def case_attention_small():
    for nq, nk in ((32, 32), (32, 14), (64, 64)):
        B, heads = 5, 4
        q, k, v = rnd(B * nq, 256), rnd(B * nk, 256), rnd(B * nk, 256)
        mask = (torch.rand(B, nk, device=dev, generator=g) < 0.7).float()
        mask[1] = 0.0
        out = ops.attention(q, k, v, batch=B, num_heads=heads, nq=nq, nk=nk, key_mask=mask)
        close(out, _attn_ref(q, k, v, mask, B, heads, nq, nk), 0.03, f"attention {nq}x{nk}")


def case_attention_tc():
    B, heads, nq, nk = 3, 4, 64, 448
    q, k, v = rnd(B * nq, 256), rnd(B * nk, 256), rnd(B * nk, 256)
    mask = (torch.arange(nk, device=dev)[None, :] < torch.tensor([448, 0, 130], device=dev)[:, None]).float()
    out = ops.attention(q, k, v, batch=B, num_heads=heads, nq=nq, nk=nk, key_mask=mask)
    close(out, _attn_ref(q, k, v, mask, B, heads, nq, nk), 0.03, "attention_tc 64x448")


def case_kv_attention():
    B, S, heads, E = 5, 192, 2, 128
    H = heads * 64
    x, wk, wv = rnd(B * S, E), rnd(H, E, scale=E ** -0.5), rnd(H, E, scale=E ** -0.5)
    bv, q = rnd(H, scale=0.5, dtype=torch.float32), rnd(B * 64, H)
    mask = (torch.arange(S, device=dev)[None, :] < torch.tensor([192, 0, 1, 100, 64], device=dev)[:, None]).float()
    out = ops.kv_attention(x, ops.pack_kv_weights(wk, wv), q, bv, batch=B, num_heads=heads, nk=S, key_mask=mask)
    k = (x.float() @ wk.float().t()).to(bf)
    v = (x.float() @ wv.float().t() + bv).to(bf)
    close(out, _attn_ref(q, k, v, mask, B, heads, 64, S), 0.04, "kv_attention 64x192")


def case_score_topk():
    for B, N, k in ((130, 70001, 100), (5, 50, 10), (64, 4096, 100)):
        u, c = rnd(B, 128), rnd(N, 128)
        s, i = ops.score_topk(u, c, k)
        full = torch.nn.functional.normalize(u.float(), dim=-1) @ torch.nn.functional.normalize(c.float(), dim=-1).t()
        rs, _ = torch.topk(full, min(k, N), dim=-1)
        close(s[:, :min(k, N)], rs, 2e-5, f"score_topk {B}x{N}")
    # descending candidates: every tile beats the running threshold -> list compaction in the epilogue
    B, N = 128, 66000
    u = torch.ones(B, 64, device=dev).to(bf)
    c = (torch.linspace(1.0, 2.0, N, device=dev)[:, None] * torch.ones(1, 64, device=dev))
    c[:, 0] = torch.linspace(3.0, -3.0, N, device=dev)
    s, i = ops.score_topk(u, c.to(bf), 100)
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    sa, ia = rnd(4, 7, 100, dtype=torch.float32), torch.randint(0, 1 << 40, (4, 7, 100), device=dev)
    ms, mi = ops.topk_merge(sa, ia)
    close(ms, torch.topk(sa.permute(1, 0, 2).reshape(7, 400), 100, dim=-1)[0], 0.0, "topk_merge")


def case_rowwise():
    x, gm, bt = rnd(300, 1024, dtype=torch.float32), rnd(1024, dtype=torch.float32), rnd(1024, dtype=torch.float32)
    close(ops.layernorm(x, gm, bt, 1e-12), torch.nn.functional.layer_norm(x, (1024,), gm, bt, 1e-12), 0.03, "layernorm fp32 in")
    xb = x.to(bf)
    close(ops.layernorm(xb, gm, bt, 1e-12), torch.nn.functional.layer_norm(xb.float(), (1024,), gm, bt, 1e-12), 0.03, "layernorm bf16 in")
    t = rnd(7, 32, 256)
    close(ops.mean_tokens(t), t.float().mean(1), 0.01, "mean_tokens")
    table = rnd(40, 32, 256, scale=0.5)
    hist = torch.randint(0, 40, (5, 7), device=dev, generator=g)
    hist[0, 1] = -1
    lens = torch.tensor([7, 1, 3, 7, 0], dtype=torch.int32, device=dev)
    seq, m = ops.build_user_sequence(table, hist, lens)
    assert seq.shape == (5, 224, 256) and float(m.sum()) == 18 * 32
    close(ops.inv_l2_norm(t.view(-1, 256)), 1.0 / t.float().view(-1, 256).norm(dim=-1), 1e-3, "inv_l2_norm")


def case_train_kernels():
    M, H = 520, 1024
    x, dy, dy2 = rnd(M, H), rnd(M, H, scale=0.1), rnd(M, H, scale=0.1)
    gm = rnd(H, dtype=torch.float32)
    dg, db = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    dx = ops.layernorm_backward(x, dy, gm, 1e-12, dg, db, dy2=dy2)
    xr = x.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (H,), gm, torch.zeros_like(gm), 1e-12).backward(dy.float() + dy2.float())
    close(dx, xr.grad, 0.02, "layernorm_backward dx")
    z = rnd(M, H)
    close(ops.gelu(z), torch.nn.functional.gelu(z.float()), 0.02, "gelu")
    close(ops.gelu_backward(z, dy), torch.autograd.grad(torch.nn.functional.gelu(zr := z.float().requires_grad_(True)).sum(), zr)[0] * dy.float(), 0.02, "gelu_backward")
    B, heads, nq, nk = 4, 4, 32, 14
    q, k, v, do = rnd(B * nq, 256), rnd(B * nk, 256), rnd(B * nk, 256), rnd(B * nq, 256, scale=0.1)
    dq, dk, dv = torch.empty_like(q), torch.zeros_like(k), torch.zeros_like(v)
    ops.attention_backward(q, k, v, do, dq, dk, dv, batch=B, num_heads=heads, nq=nq, nk=nk)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    _attn_ref(qr, kr, vr, None, B, heads, nq, nk).backward(do.float())
    close(dq, qr.grad, 0.02, "attention_backward dq")
    close(dv, vr.grad, 0.02, "attention_backward dv")
    drop = (ops.dropout_threshold(0.2), 1234, 3)
    y = ops.dropout_add(x, dy, drop)
    assert bool(torch.isfinite(y.float()).all())


def case_lists():
    B, C, D = 9, 11, 256
    u, p, c = rnd(B, D, dtype=torch.float32), rnd(B, D, dtype=torch.float32), rnd(B, C, D, dtype=torch.float32)
    mask = (torch.rand(B, C, device=dev, generator=g) < 0.8)
    sims, inv = ops.list_scores(u, p, c, mask=mask)
    ref0 = torch.nn.functional.cosine_similarity(u, p, dim=-1)
    close(sims[:, 0], ref0, 1e-4, "list_scores positive")
    loss, rank = ops.infonce_rank(sims, 0.05)
    assert bool(torch.isfinite(loss).all())


CASES = {n[5:]: f for n, f in list(globals().items()) if n.startswith("case_")}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        print(f"[{n}]", flush=True)
        CASES[n]()
        torch.cuda.synchronize()
    print("sanitizer cases ok:", " ".join(names))
