"""GPU parity of the fused K/V-projection + cross-attention kernel (csrc/kv_attention_fused.cu, C ABI
unirec_kv_attention_fused) against (a) torch fp32 of the reference's arithmetic - key / value Linear of the encoder states,
scaled dot product + additive mask, softmax, P V (models/qformer.py:185-188, 205, 244-268) - on the same bf16-rounded
inputs, and (b) the materialised path of this repo (projection GEMM + attention kernels), which stays in as the checker.

Tolerance: bf16 K / V / P with fp32 accumulation and fp32 softmax statistics - max|d| <= 0.03 on context values of
magnitude ~1, cosine >= 0.9995 (the materialised path meets the same bar)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True, params=["umma", "mma_sync"])
def kv_attention_impl(request, monkeypatch):
    """Every test of this file runs on both variants of the fused kernel (UNIREC_KV_ATTENTION_IMPL is read per call):
    "umma" - S / PV on tcgen05 with the attention state in registers, four partials per (user, head); "mma_sync" - the
    first version (attention on mma.sync over the shared K/V tile, two partials)."""
    monkeypatch.setenv("UNIREC_KV_ATTENTION_IMPL", request.param)
    return request.param


def _reference(x, wk, bk, wv, bv, q, mask, B, S, heads, q_broadcast):
    """fp32 torch restatement (mask: finfo.min added to masked keys, all-masked rows come out uniform)."""
    H = heads * 64
    k = (x.float() @ wk.float().t() + bk).view(B, S, heads, 64).transpose(1, 2)
    v = (x.float() @ wv.float().t() + bv).view(B, S, heads, 64).transpose(1, 2)
    qq = q.float().view(1 if q_broadcast else B, 64, heads, 64).transpose(1, 2)
    sc = qq @ k.transpose(-1, -2) / 8.0
    if mask is not None:
        sc = sc + ((1.0 - mask) * torch.finfo(torch.float32).min)[:, None, None, :]
    return (torch.softmax(sc, dim=-1) @ v).transpose(1, 2).reshape(B * 64, H)


def _case(B, S, heads, E, seed, masked=True, q_broadcast=False, sharp=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    H = heads * 64
    x = torch.randn(B * S, E, device=DEV, generator=g).to(torch.bfloat16)
    wk = (torch.randn(H, E, device=DEV, generator=g) * (sharp / E ** 0.5)).to(torch.bfloat16)
    wv = (torch.randn(H, E, device=DEV, generator=g) / E ** 0.5).to(torch.bfloat16)
    bk = torch.randn(H, device=DEV, generator=g) * 0.5
    bv = torch.randn(H, device=DEV, generator=g) * 0.5
    q = (torch.randn(64 if q_broadcast else B * 64, H, device=DEV, generator=g) * sharp).to(torch.bfloat16)
    mask = None
    if masked:
        lens = torch.randint(1, S + 1, (B,), device=DEV, generator=g)
        lens[0] = S
        if B > 2:
            lens[1] = 0                       # an all-masked user: uniform attention over all S keys
            lens[2] = 1
        mask = (torch.arange(S, device=DEV)[None, :] < lens[:, None]).float()
        if B > 3:
            mask[3] = (torch.rand(S, device=DEV, generator=g) < 0.5).float()      # holes, not a prefix
            mask[3, 5] = 1.0
    return x, wk, bk, wv, bv, q, mask


def _check(got, ref, what, max_tol=0.03, cos_tol=0.9995):
    got, ref = got.float(), ref.float()
    assert bool(torch.isfinite(got).all()), what
    d = float((got - ref).abs().max())
    cos = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0))
    print(f"{what}: max|d|={d:.4f} cos={cos:.6f} |ref|max={float(ref.abs().max()):.2f}")
    assert d <= max_tol and cos >= cos_tol, (what, d, cos)
    return d


@pytest.mark.parametrize("B,S,heads,E", [
    (5, 64, 4, 256),        # one 64-key half per user: the second CTA of a pair never sees some users (empty partials)
    (7, 128, 4, 256),       # two halves per user, both in the first CTA of a tile
    (6, 192, 2, 128),       # three halves: users straddle CTAs and tiles; a single head pair
    (9, 320, 4, 256),       # S % 256 != 0: items group 4 users; 9 users = a ragged last group
    (3, 1600, 16, 1024),    # the production shape (50 items x 32 tokens, 16 heads), ragged group of 3 users
    (4, 512, 4, 256),       # S % 256 == 0: one user per item
])
def test_kv_attention_matches_fp32_reference_and_materialised_path(B, S, heads, E):
    from unirec_b200 import ops
    x, wk, bk, wv, bv, q, mask = _case(B, S, heads, E, seed=B * 1000 + S)
    H = heads * 64
    got = ops.kv_attention(x, ops.pack_kv_weights(wk, wv), q, bv, batch=B, num_heads=heads, nk=S, key_mask=mask)
    ref = _reference(x, wk, bk, wv, bv, q, mask, B, S, heads, False)
    d_fused = _check(got, ref, f"fused vs fp32 [B={B} S={S} heads={heads}]")
    kv = ops.linear(x, torch.cat([wk, wv], 0).contiguous(), torch.cat([bk, bv], 0).contiguous())
    mat = ops.attention(q, kv[:, :H], kv[:, H:], batch=B, num_heads=heads, nq=64, nk=S, key_mask=mask)
    d_mat = _check(mat, ref, "materialised vs fp32")
    # two bf16 results that are each within d of the fp32 reference: at most the sum apart, plus one output ulp (2^-7 at |x| < 4)
    _check(got, mat, "fused vs materialised", max_tol=d_fused + d_mat + 2.0 ** -7)


def test_kv_attention_shared_queries_no_mask_and_sharp_softmax():
    """q_broadcast (the hoisted layer 0 of the encoder: one set of queries for every user), no mask, and a sharp softmax
    (scores of magnitude ~30: the online-softmax rescaling across tiles and the merge of the two CTAs' partials matter)."""
    from unirec_b200 import ops
    B, S, heads, E = 6, 448, 4, 256
    x, wk, bk, wv, bv, q, _ = _case(B, S, heads, E, seed=77, masked=False, q_broadcast=True, sharp=4.0)
    got = ops.kv_attention(x, ops.pack_kv_weights(wk, wv), q, bv, batch=B, num_heads=heads, nk=S, q_broadcast=True)
    ref = _reference(x, wk, bk, wv, bv, q, None, B, S, heads, True)
    # scores of magnitude ~30 carry the bf16 rounding of K (2^-9 relative, ~0.06 absolute) into the exponent: the bound is
    # the error of the materialised path (which rounds K the same way) on the same input, not a fixed number
    H = heads * 64
    kv = ops.linear(x, torch.cat([wk, wv], 0).contiguous(), torch.cat([bk, bv], 0).contiguous())
    mat = ops.attention(q, kv[:, :H], kv[:, H:], batch=B, num_heads=heads, nq=64, nk=S, q_broadcast=True)
    d_mat = _check(mat, ref, "materialised, shared queries, sharp", max_tol=0.3, cos_tol=0.999)
    _check(got, ref, "fused, shared queries, sharp", max_tol=max(0.06, 1.5 * d_mat), cos_tol=0.999)


def test_kv_attention_key_bias_cancels_and_value_bias_adds():
    """The kernel never sees the key bias (softmax-invariant) and adds the value bias after normalisation: the reference
    with an arbitrary key bias is reproduced, and changing bv shifts the output by exactly the difference."""
    from unirec_b200 import ops
    B, S, heads, E = 4, 256, 2, 128
    x, wk, bk, wv, bv, q, mask = _case(B, S, heads, E, seed=5)
    wp = ops.pack_kv_weights(wk, wv)
    a = ops.kv_attention(x, wp, q, bv, batch=B, num_heads=heads, nk=S, key_mask=mask)
    b = ops.kv_attention(x, wp, q, None, batch=B, num_heads=heads, nk=S, key_mask=mask)
    torch.testing.assert_close(a.float() - b.float(), bv.expand(B * 64, -1), rtol=0, atol=0.035)
    ref = _reference(x, wk, bk * 7.0, wv, bv, q, mask, B, S, heads, False)
    _check(a, ref, "fused vs fp32 with a 7x key bias")


def test_kv_attention_rejects_unsupported_shapes():
    from unirec_b200 import ops
    x, wk, bk, wv, bv, q, mask = _case(2, 128, 2, 128, seed=1)
    wp = ops.pack_kv_weights(wk, wv)
    with pytest.raises(RuntimeError):
        ops.kv_attention(x[:200], wp, q, bv, batch=2, num_heads=2, nk=100)         # nk % 64 != 0
    with pytest.raises(RuntimeError):
        ops.kv_attention(x.cpu(), wp.cpu(), q.cpu(), None, batch=2, num_heads=2, nk=128)   # no CPU path
    with pytest.raises(RuntimeError):
        ops.pack_kv_weights(wk[:64], wv[:64])                                       # odd head count


def test_user_qformer_fused_kv_attention_matches_materialised_and_golden():
    """UserQFormer with `fused_kv_attention = True` against the reference golden ('full': hidden 1024, 16 heads, S = 1600
    with ragged lengths) and against the same module on the materialised path."""
    import os
    import numpy as np
    from tests.golden_cases import USER_CASES
    from unirec_b200 import synth
    from unirec_b200.modules import UserQFormer
    for name in ("full", "long"):
        c = USER_CASES[name]
        mk = c["model"]
        z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"user_{name}.npz"))
        sd = synth.user_qformer_state_dict(**mk, seed=c["seed"], attn_std=c["attn_std"])
        um = UserQFormer(hidden_size=mk["hidden"], num_hidden_layers=mk["layers"], num_attention_heads=c["heads"],
                         intermediate_size=mk["inter"], num_query_tokens=mk["num_query"], input_embedding_dim=mk["input_dim"],
                         num_item_tokens_to_predict=mk["num_predict"])
        um.load_state_dict(sd, strict=True)
        um = um.to(DEV).eval()
        xs, ms = synth.user_sequences(**c["input"])
        if xs.shape[1] % 64 != 0:
            continue
        plain = um(xs.to(DEV), ms.to(DEV))
        um.fused_kv_attention = True
        assert um.qformer.fused_kv_supported(xs.shape[1])
        fused = um(xs.to(DEV), ms.to(DEV))
        ref = torch.from_numpy(z["predicted_item_tokens"])
        for what, out in (("fused", fused), ("materialised", plain)):
            d = (out.float().cpu() - ref).abs()
            cos = float(torch.nn.functional.cosine_similarity(out.float().cpu().flatten(), ref.flatten(), dim=0))
            print(f"user[{name}] {what}: max|d|={float(d.max()):.4f} mean|d|={float(d.mean()):.5f} cos={cos:.6f}")
            assert float(d.max()) <= 0.1 and float(d.mean()) <= 0.015 and cos >= 0.9995
