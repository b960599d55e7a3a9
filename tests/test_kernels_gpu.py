"""GPU parity tests of the individual sm_100a kernels, called through the C ABI (ctypes), against
plain fp32 PyTorch restatements of the same op (floating-point kernels: tolerance stated per test)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _randn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(_dev())


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (448, 1024, 1024), (96, 3072, 1024),
                                   (1000, 4096, 1024), (300, 1024, 4096), (8192, 1024, 1024), (37, 40, 192)])
@pytest.mark.parametrize("block_n", [128, 256])
def test_linear_bias(M, N, K, block_n):
    from unirec_b200 import ops
    a = _randn(M, K, seed=1, dtype=torch.bfloat16)
    w = _randn(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    b = _randn(N, seed=3)
    out = ops.linear(a, w, b, block_n=block_n, out_dtype=torch.float32)
    ref = a.float() @ w.float().t() + b
    # fp32 accumulation of exact bf16 products: only summation-order differences
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-3)
    out_bf = ops.linear(a, w, b, block_n=block_n)
    torch.testing.assert_close(out_bf.float(), ref, rtol=1e-2, atol=2e-2)  # bf16 output rounding (2^-9)


@pytest.mark.parametrize("block_n", [128, 256])
def test_linear_gelu_and_residual(block_n):
    from unirec_b200 import ops
    M, N, K = 640, 1024, 1024
    a = _randn(M, K, seed=4, dtype=torch.bfloat16)
    w = _randn(N, K, seed=5, scale=0.04, dtype=torch.bfloat16)
    b = _randn(N, seed=6, scale=0.5)
    res = _randn(M, N, seed=7, dtype=torch.bfloat16)
    lin = a.float() @ w.float().t() + b
    out = ops.linear(a, w, b, epilogue=ops.EPI_BIAS_GELU, block_n=block_n, out_dtype=torch.float32)
    torch.testing.assert_close(out, F.gelu(lin), rtol=1e-4, atol=2e-3)  # erf approximation error 1.5e-7
    out = ops.linear(a, w, b, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, block_n=block_n,
                     out_dtype=torch.float32)
    torch.testing.assert_close(out, lin + res.float(), rtol=1e-4, atol=2e-3)
    # batch-invariant residual: row r uses residual row r % 32
    res32 = res[:32].contiguous()
    out = ops.linear(a, w, b, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res32, res_row_mod=32, block_n=block_n,
                     out_dtype=torch.float32)
    torch.testing.assert_close(out, lin + res32.float().repeat(M // 32, 1), rtol=1e-4, atol=2e-3)


def test_linear_strided_views_and_many_tiles():
    """A and out as column slices of wider buffers (fused QKV layout); more tiles than SMs so that the
    persistent loop, the smem ring and both TMEM stages wrap many times."""
    from unirec_b200 import ops
    M, N, K = 20000, 2048, 1024
    big_a = _randn(M, 2 * K, seed=8, dtype=torch.bfloat16)
    a = big_a[:, K:]
    w = _randn(N, K, seed=9, scale=0.05, dtype=torch.bfloat16)
    big_out = torch.zeros(M, 3 * N, device=_dev(), dtype=torch.bfloat16)
    ops.linear(a, w, None, out=big_out[:, N:2 * N])
    ref = a.float() @ w.float().t()
    torch.testing.assert_close(big_out[:, N:2 * N].float(), ref, rtol=1e-2, atol=2e-2)
    assert float(big_out[:, :N].abs().max()) == 0.0 and float(big_out[:, 2 * N:].abs().max()) == 0.0
    # a handful of CTAs only: every CTA loops over many tiles
    out2 = ops.linear(a, w, None, max_ctas=5)
    torch.testing.assert_close(out2.float(), ref, rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (128, 256, 128), (300, 512, 1024), (2048, 1024, 1024),
                                   (1000, 4096, 1024), (777, 1024, 4096), (40000, 1024, 1024), (19, 768, 192)])
def test_linear_cta_pair_kernel(M, N, K):
    """tcgen05 cta_group::2 kernel (block_n=2): 256 x 256 tile per SM pair, TMA-store epilogue, residual via TMA.
    bf16 output: error = fp32 summation order + one bf16 rounding (2^-9 relative)."""
    from unirec_b200 import ops
    a = _randn(M, K, seed=21, dtype=torch.bfloat16)
    w = _randn(N, K, seed=22, scale=0.05, dtype=torch.bfloat16)
    b = _randn(N, seed=23, scale=0.5)
    res = _randn(M, N, seed=24, dtype=torch.bfloat16)
    lin = a.float() @ w.float().t() + b
    out = ops.linear(a, w, b, block_n=2)
    torch.testing.assert_close(out.float(), lin, rtol=1e-2, atol=2e-2)
    out = ops.linear(a, w, None, block_n=2)
    torch.testing.assert_close(out.float(), lin - b, rtol=1e-2, atol=2e-2)
    out = ops.linear(a, w, b, epilogue=ops.EPI_BIAS_GELU, block_n=2)
    torch.testing.assert_close(out.float(), F.gelu(lin), rtol=1e-2, atol=2e-2)
    out = ops.linear(a, w, b, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, block_n=2)
    torch.testing.assert_close(out.float(), lin + res.float(), rtol=1e-2, atol=3e-2)
    # identical to the single-CTA kernel up to summation order inside the tensor core (same K order: exact)
    out1 = ops.linear(a, w, b, block_n=256)
    assert float((ops.linear(a, w, b, block_n=2).float() - out1.float()).abs().max()) <= 2e-2


def test_linear_cta_pair_strided_and_few_clusters():
    from unirec_b200 import ops
    M, N, K = 20000, 2048, 1024
    big_a = _randn(M, 2 * K, seed=8, dtype=torch.bfloat16)
    a = big_a[:, K:]
    w = _randn(N, K, seed=9, scale=0.05, dtype=torch.bfloat16)
    res_big = _randn(M, 2 * N, seed=10, dtype=torch.bfloat16)
    res = res_big[:, N:]
    big_out = torch.zeros(M, 3 * N, device=_dev(), dtype=torch.bfloat16)
    ops.linear(a, w, None, out=big_out[:, N:2 * N], block_n=2, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res)
    ref = a.float() @ w.float().t() + res.float()
    torch.testing.assert_close(big_out[:, N:2 * N].float(), ref, rtol=1e-2, atol=3e-2)
    assert float(big_out[:, :N].abs().max()) == 0.0 and float(big_out[:, 2 * N:].abs().max()) == 0.0
    out2 = ops.linear(a, w, None, block_n=2, max_ctas=6)     # 3 clusters: every pair loops over many tiles
    torch.testing.assert_close(out2.float(), a.float() @ w.float().t(), rtol=1e-2, atol=2e-2)
    with pytest.raises(RuntimeError):
        ops.linear(a, w[:1000], None, block_n=2)             # N % 256 != 0
    with pytest.raises(RuntimeError):
        ops.linear(a, w, None, block_n=2, out_dtype=torch.float32)


@pytest.mark.parametrize("M,H,I", [(300, 256, 512), (2048, 1024, 4096), (37, 1024, 1024)])
def test_linear_with_folded_layernorm(M, H, I):
    """unirec_linear_ln_bf16: the LayerNorm between two GEMMs is never materialised (models/qformer.py:285-289, 371-375).
    Producer: pre = dense(x) + residual, row statistics of the stored bf16 values.  Consumers: Linear(LN(pre)) through
    gamma-scaled weights + rank-one epilogue correction (plain and GELU), and dense(y) + LN(pre) as a residual.
    Reference: fp32 torch on the same bf16 inputs; tolerance = bf16 output rounding (2^-9 relative) + accumulated input
    rounding, rtol 1.5e-2 / atol 3e-2 on O(1) outputs."""
    from unirec_b200 import ops
    eps = 1e-12
    x = _randn(M, I, seed=11, dtype=torch.bfloat16)
    w2 = _randn(H, I, seed=12, scale=0.03, dtype=torch.bfloat16)
    b2 = _randn(H, seed=13, scale=0.3)
    res = _randn(M, H, seed=14, dtype=torch.bfloat16) + 0.25      # a non-zero row mean
    gamma = 1.0 + _randn(H, seed=15, scale=0.1)
    beta = _randn(H, seed=16, scale=0.1)
    # --- producer: pre + statistics
    stats = ops.ln_stats_buffer(M, H, _dev()).fill_(float("nan"))        # every slot must be written
    pre = ops.linear_ln(x, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, stats_out=stats, eps=eps, hidden=H)
    pre_ref = x.float() @ w2.float().t() + b2 + res.float()
    torch.testing.assert_close(pre.float(), pre_ref, rtol=1e-2, atol=2e-2)
    pf = pre.float()                                             # the statistics describe the STORED values
    P = ops.ln_stats_parts(H)                                     # one partial per epilogue warp part: H / 128 or H / 64
    assert tuple(stats.shape) == (M, P, 2) and P in (H // 128, H // 64)
    torch.testing.assert_close(stats[..., 0], pf.view(M, P, H // P).sum(2), rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(stats[..., 1], (pf * pf).view(M, P, H // P).sum(2), rtol=1e-5, atol=1e-3)
    # deterministic: a second launch writes the same bits (no atomics)
    stats_b = ops.ln_stats_buffer(M, H, _dev())
    pre_b = ops.linear_ln(x, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, stats_out=stats_b, eps=eps, hidden=H)
    assert torch.equal(pre, pre_b) and torch.equal(stats, stats_b)
    h_ref = F.layer_norm(pf, (H,), gamma, beta, eps)
    # --- consumer 1: Linear(LN(pre)), bias and GELU epilogues
    w1 = _randn(I, H, seed=17, scale=0.03)
    b1 = _randn(I, seed=18, scale=0.3)
    wf, bf, cf = ops.fold_layernorm_weights(w1, b1, gamma, beta)
    lin_ref = h_ref @ w1.t() + b1
    y = ops.linear_ln(pre, wf, bf, ln_in=(stats, cf), eps=eps, hidden=H)
    torch.testing.assert_close(y.float(), lin_ref, rtol=1.5e-2, atol=3e-2)
    y = ops.linear_ln(pre, wf, bf, epilogue=ops.EPI_BIAS_GELU, ln_in=(stats, cf), eps=eps, hidden=H)
    torch.testing.assert_close(y.float(), F.gelu(lin_ref), rtol=1.5e-2, atol=3e-2)
    # against the materialised pair of kernels (LayerNorm kernel + plain GEMM): same error class
    h = ops.layernorm(pre, gamma, beta, eps)
    y_mat = ops.linear(h, w1.to(torch.bfloat16), b1, epilogue=ops.EPI_BIAS_GELU)
    e_fold = float((y.float() - F.gelu(lin_ref)).abs().max())
    e_mat = float((y_mat.float() - F.gelu(lin_ref)).abs().max())
    print(f"folded LN->GEMM max|d| {e_fold:.4f}, materialised {e_mat:.4f}")
    assert e_fold <= max(2.0 * e_mat, 2e-2)
    # --- consumer 2: dense(inter) + LN(pre) as the residual, statistics of the result
    inter = _randn(M, I, seed=19, dtype=torch.bfloat16)
    stats2 = ops.ln_stats_buffer(M, H, _dev())
    pre2 = ops.linear_ln(inter, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, residual=pre, ln_res=(stats, gamma, beta),
                         stats_out=stats2, eps=eps, hidden=H)
    pre2_ref = inter.float() @ w2.float().t() + b2 + h_ref
    torch.testing.assert_close(pre2.float(), pre2_ref, rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(stats2[..., 0].sum(1), pre2.float().sum(1), rtol=1e-5, atol=2e-3)
    # all three at once: A and the residual are the same folded tensor (the FFN-down GEMM never has that, the kernel allows it)
    if I == H:
        y3 = ops.linear_ln(pre, wf, bf, epilogue=ops.EPI_BIAS_RESIDUAL, residual=pre, ln_in=(stats, cf),
                           ln_res=(stats, gamma, beta), eps=eps, hidden=H)
        torch.testing.assert_close(y3.float(), lin_ref + h_ref, rtol=1.5e-2, atol=3e-2)


def test_linear_ln_rejects_bad_arguments():
    from unirec_b200 import ops
    a = _randn(64, 256, seed=1, dtype=torch.bfloat16)
    w = _randn(256, 256, seed=2, dtype=torch.bfloat16)
    b = _randn(256, seed=3)
    st = ops.ln_stats_buffer(64, 256, _dev())
    with pytest.raises(RuntimeError):          # N not a multiple of 256
        ops.linear_ln(a, w[:128].contiguous(), b[:128].contiguous(), stats_out=st)
    with pytest.raises(RuntimeError):          # statistics of the wrong shape
        ops.linear_ln(a, w, b, stats_out=torch.zeros(64, 2, device=_dev()))
    with pytest.raises(RuntimeError):          # residual statistics without the residual epilogue
        ops.linear_ln(a, w, b, ln_res=(st, b, b))
    with pytest.raises(RuntimeError):          # CPU tensors
        ops.linear_ln(a.cpu(), w.cpu(), b.cpu())


def test_linear_rejects_bad_arguments():
    from unirec_b200 import ops
    a = _randn(64, 100, dtype=torch.bfloat16)  # K not a multiple of 64
    w = _randn(64, 100, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.linear(a, w, None)
    with pytest.raises(RuntimeError):
        ops.linear(a.cpu(), w, None)  # no CPU path


# ------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("H", [256, 1024, 4096])
@pytest.mark.parametrize("in_dtype", [torch.float32, torch.bfloat16])
def test_layernorm(H, in_dtype):
    from unirec_b200 import ops
    rows = 333
    x = _randn(rows, H, seed=10, scale=3.0, dtype=in_dtype)
    res = _randn(rows, H, seed=11, dtype=torch.bfloat16)
    g = _randn(H, seed=12, scale=0.1) + 1.0
    b = _randn(H, seed=13, scale=0.1)
    out = ops.layernorm(x, g, b, 1e-12, residual=res, out_dtype=torch.float32)
    ref = F.layer_norm(x.float() + res.float(), (H,), g, b, 1e-12)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    out_b = ops.layernorm(x, g, b, 1e-12, rows=96, in_row_mod=32, out_dtype=torch.float32)
    ref_b = F.layer_norm(x[:32].float(), (H,), g, b, 1e-12).repeat(3, 1)
    torch.testing.assert_close(out_b, ref_b, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("H", [256, 1024])
def test_layernorm_streaming_fast_path(H):
    """bf16 in, no residual, >= 4096 rows: persistent kernel with register prefetch of the next row."""
    from unirec_b200 import ops
    rows = 5003
    wide = _randn(rows, H + 64, seed=14, scale=2.0, dtype=torch.bfloat16)
    x = wide[:, :H]                                   # strided rows
    g = _randn(H, seed=15, scale=0.1) + 1.0
    b = _randn(H, seed=16, scale=0.1)
    ref = F.layer_norm(x.float(), (H,), g, b, 1e-12)
    out = ops.layernorm(x, g, b, 1e-12, out_dtype=torch.float32)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    out_b = ops.layernorm(x, g, b, 1e-12)
    torch.testing.assert_close(out_b.float(), ref, rtol=1e-2, atol=2e-2)


# ------------------------------------------------------------------------------------- attention
def _ref_attention(q, k, v, mask, heads):
    B, nq, hd = q.shape
    nk = k.shape[1]
    qh = q.float().view(B, nq, heads, 64).permute(0, 2, 1, 3)
    kh = k.float().view(B, nk, heads, 64).permute(0, 2, 1, 3)
    vh = v.float().view(B, nk, heads, 64).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / 8.0
    if mask is not None:
        s = s + (1.0 - mask[:, None, None, :]) * torch.finfo(torch.float32).min
    p = torch.softmax(s, dim=-1)
    return (p @ vh).permute(0, 2, 1, 3).reshape(B, nq, hd)


@pytest.mark.parametrize("nq,nk,heads", [(32, 32, 16), (32, 14, 16), (64, 64, 16), (64, 160, 16), (64, 1600, 4),
                                         (32, 6, 4), (64, 97, 4), (64, 128, 2), (40, 300, 6), (64, 129, 2)])
def test_attention(nq, nk, heads):
    from unirec_b200 import ops
    B = 5
    hd = heads * 64
    q = _randn(B, nq, hd, seed=20, dtype=torch.bfloat16)
    k = _randn(B, nk, hd, seed=21, dtype=torch.bfloat16)
    v = _randn(B, nk, hd, seed=22, dtype=torch.bfloat16)
    mask = (torch.rand(B, nk, generator=torch.Generator().manual_seed(23)) < 0.7).float()
    mask[0] = 1.0
    mask[1] = 0.0          # all keys masked -> uniform attention over all nk keys
    mask[2, : nk // 2] = 0.0
    mask = mask.to(_dev())
    for m in (None, mask):
        out = ops.attention(q.view(B * nq, hd), k.view(B * nk, hd), v.view(B * nk, hd), batch=B, num_heads=heads,
                            nq=nq, nk=nk, key_mask=m).view(B, nq, hd)
        ref = _ref_attention(q, k, v, m, heads)
        assert torch.isfinite(out).all()
        # P is rounded to bf16 before PV and the output is bf16: |err| <~ 2^-8 * |v|max
        torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
    if True:
        uniform = v[1].float().mean(dim=0, keepdim=True).expand(nq, hd)
        torch.testing.assert_close(out[1].float(), uniform, rtol=2e-2, atol=2e-2)


def test_attention_two_group_kernel_in_subprocess():
    """attention_pp.cu (two softmax groups on alternate key tiles; optional, UNIREC_ATTENTION_PP=1 is read once per process):
    the long-key attention tests of this file in a child process that has the kernel switched on."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, UNIREC_ATTENTION_PP="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_kernels_gpu.py"), "-m", "gpu", "-q",
                        "-x", "-p", "no:cacheprovider", "--timeout", "120", "-k",
                        "test_attention and not two_group and not torch_ops"], env=env, cwd=root, capture_output=True,
                       text=True, timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_attention_long_keys_growing_scores_and_strided_kv():
    """tcgen05 path (nk > 64): K/V as column slices of a wide [rows, 8*H] buffer (the user model's kv_all layout),
    many work items per CTA, and scores that grow tile after tile so that the lazily raised running max and the
    rescale of the TMEM accumulator are exercised (each key tile beats the previous by > 2^8)."""
    from unirec_b200 import ops
    B, nq, nk, heads = 40, 64, 700, 16
    hd = heads * 64
    g = torch.Generator().manual_seed(60)
    q = torch.randn(B, nq, hd, generator=g)
    k = torch.randn(B, nk, hd, generator=g) * 0.3
    # add a component along the mean query direction of each head that grows with the key index
    qdir = q.view(B, nq, heads, 64).mean(dim=1)                                  # [B, heads, 64]
    qdir = qdir / qdir.norm(dim=-1, keepdim=True)
    ramp = (torch.arange(nk).float() / 128.0).floor() * 3.0                      # +3 per 128-key tile
    k = k + (ramp[None, :, None, None] * qdir[:, None, :, :]).reshape(B, nk, hd)
    v = torch.randn(B, nk, hd, generator=g)
    q, k, v = q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
    wide = torch.zeros(B * nk, 8 * hd, dtype=torch.bfloat16)
    wide[:, 2 * hd:3 * hd] = k.view(B * nk, hd)
    wide[:, 3 * hd:4 * hd] = v.view(B * nk, hd)
    wide = wide.to(_dev())
    mask = torch.ones(B, nk)
    mask[3, 500:] = 0.0
    mask[4, :650] = 0.0
    mask = mask.to(_dev())
    out = ops.attention(q.to(_dev()).view(B * nq, hd), wide[:, 2 * hd:3 * hd], wide[:, 3 * hd:4 * hd], batch=B,
                        num_heads=heads, nq=nq, nk=nk, key_mask=mask).view(B, nq, hd)
    ref = _ref_attention(q.to(_dev()), k.to(_dev()), v.to(_dev()), mask, heads)
    assert torch.isfinite(out).all()
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


def test_attention_fused_qkv_layout_and_broadcast_queries():
    """q/k/v as column slices of one fused [rows, 3*H] buffer; one query block shared by all items."""
    from unirec_b200 import ops
    B, nq, heads = 7, 32, 16
    hd = heads * 64
    qkv = _randn(B * nq, 3 * hd, seed=30, dtype=torch.bfloat16)
    q, k, v = qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:]
    out = ops.attention(q, k, v, batch=B, num_heads=heads, nq=nq, nk=nq).view(B, nq, hd)
    ref = _ref_attention(q.reshape(B, nq, hd), k.reshape(B, nq, hd), v.reshape(B, nq, hd), None, heads)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
    q1 = q[:nq].contiguous()
    out = ops.attention(q1, k, v, batch=B, num_heads=heads, nq=nq, nk=nq, q_broadcast=True).view(B, nq, hd)
    ref = _ref_attention(q1.view(1, nq, hd).expand(B, nq, hd), k.reshape(B, nq, hd), v.reshape(B, nq, hd), None, heads)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


# ------------------------------------------------------------------------------- row-wise kernels
def test_cast_mean_fieldproj_invnorm():
    from unirec_b200 import ops
    x = _randn(9, 32, 1024, seed=40)
    xb = ops.cast_bf16(x)
    assert xb.dtype == torch.bfloat16 and torch.equal(xb, x.to(torch.bfloat16))
    m = ops.mean_tokens(xb, out_dtype=torch.float32)
    torch.testing.assert_close(m, xb.float().mean(dim=1), rtol=1e-5, atol=1e-5)
    wp = _randn(14, 32, seed=41, scale=0.2)
    bp = _randn(14, seed=42)
    fp = ops.field_projection(xb, wp, bp, out_dtype=torch.float32)
    ref = torch.einsum("ft,bte->bfe", wp, xb.float()) + bp[None, :, None]
    torch.testing.assert_close(fp, ref, rtol=1e-4, atol=1e-4)
    inv = ops.inv_l2_norm(xb.view(-1, 1024))
    torch.testing.assert_close(inv, 1.0 / xb.float().view(-1, 1024).norm(dim=-1).clamp_min(1e-12), rtol=1e-5, atol=0)
    z = torch.zeros(4, 1024, device=_dev(), dtype=torch.bfloat16)
    assert float(ops.inv_l2_norm(z).max()) == pytest.approx(1e12, rel=1e-5)


def test_build_user_sequence_matches_oracle():
    from oracle import qformer_oracle as O
    from unirec_b200 import ops
    N, Q, D, B, Hmax = 40, 32, 1024, 6, 5
    table = _randn(N, Q, D, seed=50, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(51)
    history = torch.randint(0, N, (B, Hmax), generator=g)
    lengths = torch.tensor([5, 1, 3, 5, 2, 4], dtype=torch.int32)
    ctx = _randn(B, Hmax, D, seed=52, scale=0.3, dtype=torch.bfloat16)
    for c in (None, ctx):
        seq, mask = ops.build_user_sequence(table, history.to(_dev()), lengths.to(_dev()), c)
        ref_seq, ref_mask = O.build_user_sequences(table.cpu().float(), history, lengths.long(),
                                                   None if c is None else c.cpu().float())
        assert torch.equal(mask.cpu(), ref_mask)
        # bf16 output rounding of values up to ~5
        torch.testing.assert_close(seq.cpu().float(), ref_seq, rtol=1e-2, atol=2e-2)
        # the precomputed PE table and the in-kernel evaluation are the same closed form: identical outputs
        seq2, mask2 = ops.build_user_sequence(table, history.to(_dev()), lengths.to(_dev()), c, use_pe_table=False)
        assert torch.equal(seq, seq2) and torch.equal(mask, mask2)
    pe = ops.positional_encoding_table(Hmax * Q, D, _dev())
    torch.testing.assert_close(pe.cpu(), O.positional_encoding_table(Hmax * Q, D), rtol=0, atol=2e-4)


def test_torch_ops_match_direct_wrappers():
    """torch.ops.unirec_b200.* (unirec_b200/torch_ops.py) launch the same kernels as the ctypes wrappers the modules
    call directly: identical bits, and FakeTensorMode infers the real outputs' shapes and dtypes."""
    import unirec_b200.torch_ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    from unirec_b200 import ops
    ns = torch.ops.unirec_b200
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(300, 1024, generator=g).to(torch.bfloat16).to(_dev())
    w = (torch.randn(512, 1024, generator=g) * 0.03).to(torch.bfloat16).to(_dev())
    b = torch.randn(512, generator=g).to(_dev())
    assert torch.equal(ns.linear(x, w, b, None, 1, False), ops.linear(x, w, b, epilogue=ops.EPI_BIAS_GELU))
    gam, bet = torch.rand(1024, generator=g).to(_dev()), torch.randn(1024, generator=g).to(_dev())
    assert torch.equal(ns.layernorm(x, gam, bet, 1e-12, None, 0, 0, False), ops.layernorm(x, gam, bet, 1e-12))
    # LayerNorm folding: producer (statistics out) and consumer (statistics in) through the dispatcher = the direct wrappers
    res = torch.randn(300, 512, generator=g).to(torch.bfloat16).to(_dev())
    st = ops.ln_stats_buffer(300, 512, _dev())
    pre = ops.linear_ln(x, w, b, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res, stats_out=st, hidden=512)
    pre_t, st_t = ns.linear_ln(x, w, b, res, 2, None, None, None, None, None, True, 1e-12, 512)
    assert torch.equal(pre, pre_t) and torch.equal(st, st_t)
    w1 = (torch.randn(256, 512, generator=g) * 0.03).to(_dev())
    wf, bf_, cf = ops.fold_layernorm_weights(w1, None, gam[:512], bet[:512])
    y_t, none_t = ns.linear_ln(pre, wf, bf_, None, 1, st, cf, None, None, None, False, 1e-12, 512)
    assert torch.equal(y_t, ops.linear_ln(pre, wf, bf_, epilogue=ops.EPI_BIAS_GELU, ln_in=(st, cf), hidden=512))
    assert none_t.numel() == 0
    B, heads, nq, nk = 3, 16, 32, 14
    q = torch.randn(B * nq, 1024, generator=g).to(torch.bfloat16).to(_dev())
    k = torch.randn(B * nk, 1024, generator=g).to(torch.bfloat16).to(_dev())
    v = torch.randn(B * nk, 1024, generator=g).to(torch.bfloat16).to(_dev())
    mask = (torch.rand(B, nk, generator=g) < 0.8).float().to(_dev())
    assert torch.equal(ns.attention(q, k, v, mask, B, heads, nq, nk, False),
                       ops.attention(q, k, v, batch=B, num_heads=heads, nq=nq, nk=nk, key_mask=mask))
    cands = torch.randn(5000, 1024, generator=g).to(torch.bfloat16).to(_dev())
    users = torch.randn(6, 1024, generator=g).to(torch.bfloat16).to(_dev())
    s1, i1 = ns.score_topk(users, cands, 10, None, None, 7)
    s2, i2 = ops.score_topk(users, cands, 10, index_base=7)
    assert torch.equal(s1, s2) and torch.equal(i1, i2)
    with FakeTensorMode(allow_non_fake_inputs=False) as mode:
        fx, fw, fb = mode.from_tensor(x), mode.from_tensor(w), mode.from_tensor(b)
        fy = ns.linear(fx, fw, fb, None, 0, True)
        assert tuple(fy.shape) == (300, 512) and fy.dtype == torch.float32 and fy.device.type == "cuda"
    with pytest.raises(NotImplementedError):
        ns.mean_tokens(torch.zeros(2, 4, 64, dtype=torch.bfloat16), False)


def test_torch_ops_autograd():
    """`register_autograd` of torch.ops.unirec_b200.{linear, layernorm, attention}: a small attention block written
    against the dispatcher ops trains under eager autograd; gradients against torch fp32 autograd of the same math on
    the same (bf16-rounded) inputs.  Tolerance: bf16 gradients - cosine >= 0.999, relative error <= 3 %."""
    import unirec_b200.torch_ops  # noqa: F401
    ns = torch.ops.unirec_b200
    g = torch.Generator(device="cpu").manual_seed(9)
    B, heads, nq, nk, H = 5, 4, 32, 14, 256

    def leaf(*shape, scale=1.0, dtype=torch.bfloat16):
        return (torch.randn(*shape, generator=g) * scale).to(dtype).to(_dev()).requires_grad_(True)

    x, enc = leaf(B * nq, H), leaf(B * nk, H)
    wq, wk, wv, wo = (leaf(H, H, scale=0.06) for _ in range(4))
    bq, bo = leaf(H, dtype=torch.float32), leaf(H, dtype=torch.float32)
    gam, bet = leaf(H, dtype=torch.float32), leaf(H, dtype=torch.float32)
    w1 = leaf(2 * H, H, scale=0.06)
    b1 = leaf(2 * H, dtype=torch.float32)
    mask = (torch.rand(B, nk, generator=g) < 0.8).float().to(_dev())
    wsum = torch.randn(B * nq, 2 * H, generator=g).to(_dev())
    params = [x, enc, wq, wk, wv, wo, bq, bo, gam, bet, w1, b1]

    q = ns.linear(x, wq, bq, None, 0, False)
    k = ns.linear(enc, wk, None, None, 0, False)
    v = ns.linear(enc, wv, None, None, 0, False)
    ctx = ns.attention(q, k, v, mask, B, heads, nq, nk, False)
    pre = ns.linear(ctx, wo, bo, x, 2, False)                       # bias + residual epilogue
    h = ns.layernorm(pre, gam, bet, 1e-12, None, 0, 0, False)
    y = ns.linear(h, w1, b1, None, 1, True)                         # bias + erf-GELU epilogue, fp32 out
    (y * wsum).sum().backward()
    got = [p.grad.float().cpu() for p in params]

    ref_p = [p.detach().float().cpu().requires_grad_(True) for p in params]
    rx, renc, rwq, rwk, rwv, rwo, rbq, rbo, rgam, rbet, rw1, rb1 = ref_p
    F = torch.nn.functional
    rq = F.linear(rx, rwq, rbq).view(B, nq, heads, 64).transpose(1, 2)
    rk = F.linear(renc, rwk).view(B, nk, heads, 64).transpose(1, 2)
    rv = F.linear(renc, rwv).view(B, nk, heads, 64).transpose(1, 2)
    sc = rq @ rk.transpose(-1, -2) / 8.0 + ((1.0 - mask.cpu()) * torch.finfo(torch.float32).min)[:, None, None, :]
    rctx = (torch.softmax(sc, dim=-1) @ rv).transpose(1, 2).reshape(B * nq, H)
    rh = F.layer_norm(F.linear(rctx, rwo, rbo) + rx, (H,), rgam, rbet, 1e-12)
    ry = F.gelu(F.linear(rh, rw1, rb1))
    (ry * wsum.cpu()).sum().backward()
    names = ["x", "enc", "wq", "wk", "wv", "wo", "bq", "bo", "gamma", "beta", "w1", "b1"]
    for n, a, r in zip(names, got, ref_p):
        cos = float(F.cosine_similarity(a.flatten(), r.grad.flatten(), dim=0))
        rel = float((a - r.grad).norm() / (r.grad.norm() + 1e-12))
        print(f"{n:6s} cos={cos:.5f} rel={rel:.4f}")
        assert cos >= 0.999 and rel <= 0.03, (n, cos, rel)
