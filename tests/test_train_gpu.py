"""GPU parity of the backward-pass kernels (C ABI "Backward pass" block) against torch fp32 autograd of the same
op on the same bf16 inputs, and of one full training step against autograd through the CPU oracle."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _randn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (304, 512, 1000), (1024, 1024, 4096), (64, 264, 136), (4096, 3072, 1024)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_general_layouts(M, N, K, a_mn, b_mn):
    from unirec_b200 import ops
    a = _randn(M, K, seed=1, dtype=torch.bfloat16)
    b = _randn(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    ref = a.float() @ b.float().t()
    a_st = a.t().contiguous() if a_mn else a
    b_st = b.t().contiguous() if b_mn else b
    out = ops.gemm_general(a_st, b_st, a_mn=a_mn, b_mn=b_mn, M=M, N=N, K=K, out_dtype=torch.float32)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=3e-3)
    out_b = ops.gemm_general(a_st, b_st, a_mn=a_mn, b_mn=b_mn, M=M, N=N, K=K)
    torch.testing.assert_close(out_b.float(), ref, rtol=1e-2, atol=3e-2)


@pytest.mark.parametrize("ksplit", [0, 1, 3, 7])
def test_gemm_general_accumulate_split_k(ksplit):
    """wgrad shape: few output tiles, long contraction, split over CTAs with fp32 atomic accumulation."""
    from unirec_b200 import ops
    rows, N, Kin = 9000 + 8, 1024, 1024
    dy = _randn(rows, N, seed=3, scale=0.1, dtype=torch.bfloat16)
    x = _randn(rows, Kin, seed=4, dtype=torch.bfloat16)
    dw = torch.full((N, Kin), 0.5, device=DEV, dtype=torch.float32)
    ops.gemm_general(dy, x, a_mn=True, b_mn=True, M=N, N=Kin, K=rows, out=dw, accumulate=True, ksplit=ksplit)
    ref = 0.5 + dy.float().t() @ x.float()
    torch.testing.assert_close(dw, ref, rtol=1e-4, atol=2e-2)


def test_linear_dgrad_wgrad_colsum_match_autograd():
    from unirec_b200 import ops
    M, N, Kin = 2048 + 32, 3072, 1024
    x = _randn(M, Kin, seed=5, dtype=torch.bfloat16)
    w = _randn(N, Kin, seed=6, scale=0.03, dtype=torch.bfloat16)
    dy = _randn(M, N, seed=7, scale=0.1, dtype=torch.bfloat16)
    xf, wf = x.float().requires_grad_(), w.float().requires_grad_()
    bf = torch.zeros(N, device=DEV, requires_grad=True)
    (F.linear(xf, wf, bf) * dy.float()).sum().backward()
    dx = ops.linear_dgrad(dy, w)
    torch.testing.assert_close(dx.float(), xf.grad, rtol=1e-2, atol=2e-2)
    dw = torch.zeros(N, Kin, device=DEV)
    ops.linear_wgrad(dy, x, dw)
    torch.testing.assert_close(dw, wf.grad, rtol=1e-4, atol=1e-2)
    db = torch.zeros(N, device=DEV)
    ops.colsum(dy, db)
    torch.testing.assert_close(db, bf.grad, rtol=1e-4, atol=1e-3)
    # strided dy (a column slice of a fused [M, 3N'] gradient buffer)
    big = _randn(M, 2 * N, seed=8, scale=0.1, dtype=torch.bfloat16)
    sl = big[:, N:]
    dx2 = ops.linear_dgrad(sl, w)
    torch.testing.assert_close(dx2.float(), sl.float() @ w.float(), rtol=1e-2, atol=2e-2)


def test_gelu_forward_backward():
    from unirec_b200 import ops
    z = _randn(777, 4096, seed=9, scale=2.0, dtype=torch.bfloat16)
    da = _randn(777, 4096, seed=10, dtype=torch.bfloat16)
    zf = z.float().requires_grad_()
    y = F.gelu(zf)
    y.backward(da.float())
    torch.testing.assert_close(ops.gelu(z).float(), y.detach(), rtol=1e-2, atol=1e-3)
    torch.testing.assert_close(ops.gelu_backward(z, da).float(), zf.grad, rtol=1e-2, atol=2e-3)


@pytest.mark.parametrize("H", [256, 1024])
def test_layernorm_backward(H):
    from unirec_b200 import ops
    rows = 3001
    x = _randn(rows, H, seed=11, scale=2.0, dtype=torch.bfloat16)
    dy = _randn(rows, H, seed=12, dtype=torch.bfloat16)
    dy2 = _randn(rows, H, seed=13, dtype=torch.bfloat16)
    g = (_randn(H, seed=14, scale=0.1) + 1.0)
    b = _randn(H, seed=15, scale=0.1)
    for extra in (None, dy2):
        xf = x.float().requires_grad_()
        gf, bfp = g.clone().requires_grad_(), b.clone().requires_grad_()
        F.layer_norm(xf, (H,), gf, bfp, 1e-12).backward(dy.float() + (0 if extra is None else extra.float()))
        dg = torch.zeros(H, device=DEV)
        db = torch.zeros(H, device=DEV)
        dx = ops.layernorm_backward(x, dy, g, 1e-12, dg, db, dy2=extra)
        torch.testing.assert_close(dx.float(), xf.grad, rtol=2e-2, atol=2e-2)
        torch.testing.assert_close(dg, gf.grad, rtol=1e-3, atol=5e-2)
        torch.testing.assert_close(db, bfp.grad, rtol=1e-3, atol=5e-2)


@pytest.mark.parametrize("nq,nk,heads", [(32, 32, 16), (32, 14, 16), (64, 64, 4), (32, 6, 4), (48, 40, 2)])
def test_attention_backward(nq, nk, heads):
    from unirec_b200 import ops
    B, hd = 6, heads * 64
    q = _randn(B, nq, hd, seed=20, dtype=torch.bfloat16)
    k = _randn(B, nk, hd, seed=21, dtype=torch.bfloat16)
    v = _randn(B, nk, hd, seed=22, dtype=torch.bfloat16)
    do = _randn(B, nq, hd, seed=23, dtype=torch.bfloat16)
    mask = (torch.rand(B, nk, generator=torch.Generator().manual_seed(24)) < 0.7).float()
    mask[0] = 1.0
    mask[:, 0] = 1.0
    mask = mask.to(DEV)
    for m in (None, mask):
        qf, kf, vf = (t.float().requires_grad_() for t in (q, k, v))
        qh = qf.view(B, nq, heads, 64).permute(0, 2, 1, 3)
        kh = kf.view(B, nk, heads, 64).permute(0, 2, 1, 3)
        vh = vf.view(B, nk, heads, 64).permute(0, 2, 1, 3)
        s = qh @ kh.transpose(-1, -2) / 8.0
        if m is not None:
            s = s + (1.0 - m[:, None, None, :]) * torch.finfo(torch.float32).min
        out = (torch.softmax(s, dim=-1) @ vh).permute(0, 2, 1, 3).reshape(B, nq, hd)
        out.backward(do.float())
        dq = torch.zeros(B * nq, hd, device=DEV, dtype=torch.bfloat16)
        dk = torch.zeros(B * nk, hd, device=DEV, dtype=torch.bfloat16)
        dv = torch.zeros(B * nk, hd, device=DEV, dtype=torch.bfloat16)
        ops.attention_backward(q.view(B * nq, hd), k.view(B * nk, hd), v.view(B * nk, hd), do.view(B * nq, hd), dq, dk, dv,
                               batch=B, num_heads=heads, nq=nq, nk=nk, key_mask=m)
        # P and dS are rounded to bf16 before the transposed products: |err| <~ 2^-8 of the gradient scale
        torch.testing.assert_close(dq.float().view(B, nq, hd), qf.grad, rtol=3e-2, atol=3e-2)
        torch.testing.assert_close(dk.float().view(B, nk, hd), kf.grad, rtol=3e-2, atol=3e-2)
        torch.testing.assert_close(dv.float().view(B, nk, hd), vf.grad, rtol=3e-2, atol=3e-2)


def test_dropout_rows_match_oracle_masks():
    """dropout_add / dropout_backward reproduce the CPU restatement of the Philox mask bit for bit."""
    from oracle import dropout_masks as DM
    from unirec_b200 import ops
    rows, H, p = 515, 256, 0.2
    thr = ops.dropout_threshold(p)
    seed, site = (1 << 40) + 12345, 19
    x = _randn(rows, H, seed=30, dtype=torch.bfloat16)
    res = _randn(rows, H, seed=31, dtype=torch.bfloat16)
    keep = torch.from_numpy(DM.keep_mask_rows(seed, site, rows, H, thr)).to(DEV)
    scale = DM.keep_scale(thr)
    assert 0.77 < float(keep.float().mean()) < 0.83
    out = ops.dropout_add(x, res, (thr, seed, site))
    ref = (x.float() * scale * keep + res.float()).to(torch.bfloat16)
    assert torch.equal(out, ref)
    out0 = ops.dropout_add(x, None, (thr, seed, site))
    assert torch.equal(out0, (x.float() * scale * keep).to(torch.bfloat16))
    dx = ops.dropout_backward(x, (thr, seed, site))
    assert torch.equal(dx, (x.float() * scale * keep).to(torch.bfloat16))
    # broadcast input (query embeddings: 32 rows serve every batch element, each with its own mask)
    q = _randn(32, H, seed=32, dtype=torch.bfloat16)
    outq = ops.dropout_add(q, None, (thr, seed, 0), rows=rows - 3, x_row_mod=32)
    keepq = torch.from_numpy(DM.keep_mask_rows(seed, 0, rows - 3, H, thr)).to(DEV)
    refq = (q.float()[torch.arange(rows - 3, device=DEV) % 32] * scale * keepq).to(torch.bfloat16)
    assert torch.equal(outq, refq)


@pytest.mark.parametrize("nq,nk,heads", [(32, 32, 16), (32, 14, 16), (64, 64, 4), (48, 40, 2)])
def test_attention_dropout_forward_backward(nq, nk, heads):
    """Probability dropout inside the attention kernels vs torch autograd with the oracle's mask."""
    from oracle import dropout_masks as DM
    from unirec_b200 import ops
    B, hd, p = 5, heads * 64, 0.2
    thr, seed, site = ops.dropout_threshold(p), 987654321012345, 1 + 8 * 3 + 2
    q = _randn(B, nq, hd, seed=40, dtype=torch.bfloat16)
    k = _randn(B, nk, hd, seed=41, dtype=torch.bfloat16)
    v = _randn(B, nk, hd, seed=42, dtype=torch.bfloat16)
    do = _randn(B, nq, hd, seed=43, dtype=torch.bfloat16)
    mask = (torch.rand(B, nk, generator=torch.Generator().manual_seed(44)) < 0.7).float()
    mask[:, 0] = 1.0
    mask = mask.to(DEV)
    keep = torch.from_numpy(DM.keep_mask_attention(seed, site, B, heads, nq, nk, thr)).to(DEV).float() * DM.keep_scale(thr)
    qf, kf, vf = (t.float().requires_grad_() for t in (q, k, v))
    qh = qf.view(B, nq, heads, 64).permute(0, 2, 1, 3)
    kh = kf.view(B, nk, heads, 64).permute(0, 2, 1, 3)
    vh = vf.view(B, nk, heads, 64).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) / 8.0 + (1.0 - mask[:, None, None, :]) * torch.finfo(torch.float32).min
    ref = ((torch.softmax(s, dim=-1) * keep) @ vh).permute(0, 2, 1, 3).reshape(B, nq, hd)
    ref.backward(do.float())
    out = ops.attention(q.view(B * nq, hd), k.view(B * nk, hd), v.view(B * nk, hd), batch=B, num_heads=heads, nq=nq, nk=nk,
                        key_mask=mask, dropout=(thr, seed, site))
    torch.testing.assert_close(out.float().view(B, nq, hd), ref.detach(), rtol=2e-2, atol=2e-2)
    dq = torch.zeros(B * nq, hd, device=DEV, dtype=torch.bfloat16)
    dk = torch.zeros(B * nk, hd, device=DEV, dtype=torch.bfloat16)
    dv = torch.zeros(B * nk, hd, device=DEV, dtype=torch.bfloat16)
    ops.attention_backward(q.view(B * nq, hd), k.view(B * nk, hd), v.view(B * nk, hd), do.view(B * nq, hd), dq, dk, dv,
                           batch=B, num_heads=heads, nq=nq, nk=nk, key_mask=mask, dropout=(thr, seed, site))
    torch.testing.assert_close(dq.float().view(B, nq, hd), qf.grad, rtol=3e-2, atol=4e-2)
    torch.testing.assert_close(dk.float().view(B, nk, hd), kf.grad, rtol=3e-2, atol=4e-2)
    torch.testing.assert_close(dv.float().view(B, nk, hd), vf.grad, rtol=3e-2, atol=4e-2)


def test_attention_dropout_long_keys_forward():
    """Key-tiled path (nk > 64: several 64-key tiles with the online softmax) with dropout."""
    from oracle import dropout_masks as DM
    from unirec_b200 import ops
    B, heads, nq, nk = 3, 4, 64, 200
    hd = heads * 64
    thr, seed, site = ops.dropout_threshold(0.1), 77, 11
    q = _randn(B, nq, hd, seed=50, dtype=torch.bfloat16)
    k = _randn(B, nk, hd, seed=51, dtype=torch.bfloat16)
    v = _randn(B, nk, hd, seed=52, dtype=torch.bfloat16)
    keep = torch.from_numpy(DM.keep_mask_attention(seed, site, B, heads, nq, nk, thr)).to(DEV).float() * DM.keep_scale(thr)
    qh = q.float().view(B, nq, heads, 64).permute(0, 2, 1, 3)
    kh = k.float().view(B, nk, heads, 64).permute(0, 2, 1, 3)
    vh = v.float().view(B, nk, heads, 64).permute(0, 2, 1, 3)
    ref = ((torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, dim=-1) * keep) @ vh).permute(0, 2, 1, 3).reshape(B, nq, hd)
    out = ops.attention(q.view(B * nq, hd), k.view(B * nk, hd), v.view(B * nk, hd), batch=B, num_heads=heads, nq=nq, nk=nk,
                        dropout=(thr, seed, site))
    torch.testing.assert_close(out.float().view(B, nq, hd), ref, rtol=2e-2, atol=2e-2)


def _oracle_loss_and_grads(sd, x, mask, heads, weights, drop=None):
    """Reference: fp32 autograd through the CPU oracle; loss = sum(outputs * fixed random weights)."""
    from oracle import qformer_oracle as O
    sd = {k: v.clone().float().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    out = O.item_qformer_forward(sd, x, mask, num_heads=heads, drop=drop)
    loss = sum((out[k] * weights[k]).sum() for k in ("query_outputs", "item_representation", "reconstructed_fields"))
    loss.backward()
    return float(loss), {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_item_training_step_gradients_match_oracle_autograd(dropout):
    """Small item Q-Former (4 layers, hidden 256): forward + backward through the CUDA training path vs torch
    autograd through the fp32 oracle on the same weights and inputs - with dropout 0 and with the reference's
    default dropout 0.2 (the oracle regenerates the CUDA path's Philox masks from the same seed)."""
    from tests.golden_cases import ITEM_CASES
    from unirec_b200 import synth
    from unirec_b200.modules import QFormerForItemRepresentation
    c = ITEM_CASES["small"]
    mk = c["model"]
    sd = synth.item_qformer_state_dict(**mk, seed=61, attn_std=0.1)
    model = QFormerForItemRepresentation(hidden_size=mk["hidden"], num_hidden_layers=mk["layers"],
                                         num_attention_heads=c["heads"], intermediate_size=mk["inter"],
                                         num_query_tokens=mk["num_query"], field_embedding_dim=mk["field_dim"],
                                         num_fields=mk["num_fields"], dropout=dropout)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).train()
    model.dropout_seed = 20261017
    B = 64
    x, mask = synth.item_fields(batch=B, num_fields=6, dim=256, seed=62, clip_field=2, presence=0.8)
    g = torch.Generator().manual_seed(63)
    weights = {"query_outputs": torch.randn(B, 32, 256, generator=g) / 100,
               "item_representation": torch.randn(B, 256, generator=g) / 10,
               "reconstructed_fields": torch.randn(B, 6, 256, generator=g) / 30}
    out = model(x.to(DEV), mask.to(DEV))
    loss = sum((out[k].float() * weights[k].to(DEV)).sum() for k in weights)
    loss.backward()
    assert (model.last_dropout is None) == (dropout == 0.0)
    ref_loss, ref = _oracle_loss_and_grads(sd, x, mask, c["heads"], weights, drop=model.last_dropout)
    # the loss is a signed random-weighted sum of ~0.6 M outputs: compare against the scale of its terms
    term_scale = float(sum((out[k].float().abs() * weights[k].to(DEV).abs()).sum() for k in weights))
    loss_err = abs(float(loss) - ref_loss)
    print(f"loss {float(loss):.4f} vs oracle {ref_loss:.4f} (sum |terms| = {term_scale:.1f})")
    checked, bad = 0, []
    for name, prm in model.named_parameters():
        if name not in ref:
            continue
        if prm.grad is None:
            assert float(ref[name].abs().max()) == 0.0, name       # dead (text-branch) tensors get no gradient
            continue
        gref = ref[name]
        got = prm.grad.float().cpu()
        if name.endswith("self.key.bias"):
            # softmax is invariant to a shift of all scores of a row, so d loss / d key.bias == 0 exactly; the
            # reference leaves fp32 noise there and so do we (bf16 noise): only require that it is small
            assert float(got.abs().max()) < 2e-2 and float(gref.abs().max()) < 1e-4, (name, float(got.abs().max()))
            continue
        cos = float(F.cosine_similarity(got.flatten(), gref.flatten(), dim=0))
        rel = float((got - gref).norm() / (gref.norm() + 1e-12))
        print(f"{name:70s} cos={cos:.5f} rel={rel:.4f} |ref|={float(gref.norm()):.3e}")
        if not (cos > 0.99 and rel < 0.12):
            bad.append((name, cos, rel))
        checked += 1
    assert not bad, bad
    assert checked > 60
    assert loss_err <= 2e-4 * term_scale + 0.02 * abs(ref_loss), (float(loss), ref_loss, term_scale)


def _compare_grads(model, ref, what, cos_tol, rel_tol):
    checked, bad, worst_cos, worst_rel = 0, [], 1.0, 0.0
    for name, prm in model.named_parameters():
        if name not in ref:
            continue
        if prm.grad is None:
            assert float(ref[name].abs().max()) == 0.0, name       # dead (text-branch) tensors get no gradient
            continue
        gref, got = ref[name], prm.grad.float().cpu()
        if name.endswith("self.key.bias"):                         # d loss / d key.bias == 0 exactly (softmax shift)
            assert float(got.abs().max()) <= 2e-2 * max(1.0, float(got.abs().max() + gref.abs().max())), name
            continue
        cos = float(F.cosine_similarity(got.flatten(), gref.flatten(), dim=0))
        rel = float((got - gref).norm() / (gref.norm() + 1e-12))
        worst_cos, worst_rel = min(worst_cos, cos), max(worst_rel, rel)
        if not (cos >= cos_tol and rel <= rel_tol):
            bad.append((name, cos, rel))
        checked += 1
    print(f"{what}: {checked} tensors, worst cos {worst_cos:.5f}, worst rel {worst_rel:.4f}")
    assert not bad, (what, bad[:8])
    return checked


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_full_size_training_step_gradients_match_oracle_eager_and_graph(dropout):
    """The BENCHED training model - 12 layers, hidden 1024, 16 heads, FFN 4096, 14 fields, B = 256 items (8192 query rows:
    the split-K wgrad contracts over 8192 / 3584 rows, layernorm_bwd runs at H = 1024 inside the chain) - with the
    reference's loss (QFormerLoss, training/item_qformer_training.py:41-56): gradients of EVERY live tensor through the
    eager path and through one replay of `TrainStepGraph` against fp32 autograd through the CPU oracle, with dropout 0
    and with the reference's dropout 0.2 (the oracle regenerates the kernels' Philox masks from the seed each path used).
    Tolerance: cosine >= 0.999 and relative L2 error <= 6 % per tensor (bf16 activations over 12 layers)."""
    import gc
    from tests.golden_cases import ITEM_CASES
    from oracle import qformer_oracle as O
    from unirec_b200 import synth
    from unirec_b200.modules import QFormerForItemRepresentation
    from unirec_b200.training import QFormerLoss, TrainStepGraph, qformer_loss
    c = ITEM_CASES["full"]
    mk = c["model"]
    sd = synth.item_qformer_state_dict(**mk, seed=71, attn_std=0.02)
    model = QFormerForItemRepresentation(hidden_size=mk["hidden"], num_hidden_layers=mk["layers"],
                                         num_attention_heads=c["heads"], intermediate_size=mk["inter"],
                                         num_query_tokens=mk["num_query"], field_embedding_dim=mk["field_dim"],
                                         num_fields=mk["num_fields"], dropout=dropout)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).train()
    model.dropout_seed = 424242
    B = 256
    x, mask = synth.item_fields(batch=B, num_fields=14, dim=1024, seed=72, presence=0.8)
    g = torch.Generator().manual_seed(73)
    pos, neg = torch.randn(B, 1024, generator=g), torch.randn(B, 1024, generator=g)
    xd, md, pd, nd = x.to(DEV), mask.to(DEV), pos.to(DEV), neg.to(DEV)

    def oracle(drop):
        leaf = {k: v.clone().float().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        out = O.item_qformer_forward(leaf, x, mask, num_heads=c["heads"], drop=drop)
        loss = QFormerLoss()(out, {"field_embeddings": x}, pos, neg, mask)[0]
        loss.backward()
        return float(loss), {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}

    # ---- eager path
    loss = qformer_loss(model(xd, md), xd, md, pd, nd)
    loss.backward()
    drop_e = model.last_dropout
    ref_loss, ref = oracle(drop_e)
    print(f"eager loss {float(loss):.4f} vs oracle {ref_loss:.4f}")
    assert abs(float(loss) - ref_loss) <= 5e-3 * abs(ref_loss)
    assert _compare_grads(model, ref, f"eager, dropout {dropout}", 0.999, 0.06) >= 240
    # ---- one replay of the captured step
    loss_e = float(loss)
    del loss
    model.zero_grad(set_to_none=True)
    gc.collect()
    tg = TrainStepGraph(model, xd, md)
    frozen = model.last_dropout                    # (thr16, seed) frozen into the captured launches
    got = float(tg.step(xd, md, pd, nd))
    if dropout > 0:
        # a replay advances the device-resident seed offset by SEED_STRIDE before the forward runs
        ref_loss, ref = oracle((frozen[0], frozen[1] + TrainStepGraph.SEED_STRIDE))
    print(f"graph loss {got:.4f} vs oracle {ref_loss:.4f} (eager {loss_e:.4f})")
    assert abs(got - ref_loss) <= 5e-3 * abs(ref_loss)
    assert _compare_grads(model, ref, f"graph, dropout {dropout}", 0.999, 0.06) >= 240


def test_train_mode_no_grad_forward_uses_dropout_and_eval_does_not():
    """The reference's step runs the positive / negative forwards under no_grad with the module in train():
    dropout stays active there (training/item_qformer_training.py:122-125); eval() is deterministic."""
    from oracle import qformer_oracle as O
    from tests.golden_cases import ITEM_CASES
    from unirec_b200 import synth
    from unirec_b200.modules import QFormerForItemRepresentation
    c = ITEM_CASES["small"]
    mk = c["model"]
    sd = synth.item_qformer_state_dict(**mk, seed=61, attn_std=0.1)
    model = QFormerForItemRepresentation(hidden_size=mk["hidden"], num_hidden_layers=mk["layers"],
                                         num_attention_heads=c["heads"], intermediate_size=mk["inter"],
                                         num_query_tokens=mk["num_query"], field_embedding_dim=mk["field_dim"],
                                         num_fields=mk["num_fields"], dropout=0.2)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).train()
    x, mask = synth.item_fields(batch=16, num_fields=6, dim=256, seed=62, clip_field=2, presence=0.8)
    model.dropout_seed = 5
    with torch.no_grad():
        a = model(x.to(DEV), mask.to(DEV))["query_outputs"].float().cpu()
        drop_a = model.last_dropout
        b = model(x.to(DEV), mask.to(DEV))["query_outputs"].float().cpu()
    assert drop_a == (13107, 5) and model.last_dropout == (13107, 6)
    assert float((a - b).abs().max()) > 0.05                       # different seeds, different masks
    ref = O.item_qformer_forward(sd, x, mask, num_heads=c["heads"], drop=drop_a)["query_outputs"]
    assert float((a - ref).abs().mean()) < 0.02 and float(F.cosine_similarity(a.flatten(), ref.flatten(), dim=0)) > 0.999
    model.eval()
    e1 = model(x.to(DEV), mask.to(DEV))["query_outputs"]
    e2 = model(x.to(DEV), mask.to(DEV))["query_outputs"]
    assert torch.equal(e1, e2)


def test_dropout_seed_offset_equals_seed_addition():
    """The device-resident seed offset (include/unirec_b200.h: seed_offset) is exactly `seed + *seed_offset`: row
    dropout and attention dropout with (seed, offset = d) reproduce (seed + d, no offset) bit for bit."""
    from unirec_b200 import ops
    thr = ops.dropout_threshold(0.2)
    seed, d, site = (1 << 33) + 77, 4096 + 5, 11
    off = torch.tensor([d], device=DEV, dtype=torch.int64)
    x = _randn(300, 256, seed=70, dtype=torch.bfloat16)
    res = _randn(300, 256, seed=71, dtype=torch.bfloat16)
    assert torch.equal(ops.dropout_add(x, res, (thr, seed, site, off)), ops.dropout_add(x, res, (thr, seed + d, site)))
    assert torch.equal(ops.dropout_backward(x, (thr, seed, site, off)), ops.dropout_backward(x, (thr, seed + d, site)))
    assert not torch.equal(ops.dropout_add(x, res, (thr, seed, site, off)), ops.dropout_add(x, res, (thr, seed, site)))
    B, heads, nq, nk = 5, 4, 32, 14
    q = _randn(B * nq, heads * 64, seed=72, dtype=torch.bfloat16)
    k = _randn(B * nk, heads * 64, seed=73, dtype=torch.bfloat16)
    v = _randn(B * nk, heads * 64, seed=74, dtype=torch.bfloat16)
    a = ops.attention(q, k, v, batch=B, num_heads=heads, nq=nq, nk=nk, dropout=(thr, seed, site, off))
    b = ops.attention(q, k, v, batch=B, num_heads=heads, nq=nq, nk=nk, dropout=(thr, seed + d, site))
    assert torch.equal(a, b)
    do = _randn(B * nq, heads * 64, seed=75, dtype=torch.bfloat16)
    g1 = [torch.empty_like(t) for t in (q, k, v)]
    g2 = [torch.empty_like(t) for t in (q, k, v)]
    ops.attention_backward(q, k, v, do, *g1, batch=B, num_heads=heads, nq=nq, nk=nk, dropout=(thr, seed, site, off))
    ops.attention_backward(q, k, v, do, *g2, batch=B, num_heads=heads, nq=nq, nk=nk, dropout=(thr, seed + d, site))
    for t1, t2 in zip(g1, g2):
        assert torch.equal(t1, t2)


def _small_train_model(dropout, seed=61):
    from tests.golden_cases import ITEM_CASES
    from unirec_b200 import synth
    from unirec_b200.modules import QFormerForItemRepresentation
    c = ITEM_CASES["small"]
    mk = c["model"]
    sd = synth.item_qformer_state_dict(**mk, seed=seed, attn_std=0.1)
    model = QFormerForItemRepresentation(hidden_size=mk["hidden"], num_hidden_layers=mk["layers"],
                                         num_attention_heads=c["heads"], intermediate_size=mk["inter"],
                                         num_query_tokens=mk["num_query"], field_embedding_dim=mk["field_dim"],
                                         num_fields=mk["num_fields"], dropout=dropout)
    model.load_state_dict(sd, strict=True)
    return model.to(DEV).train()


def test_train_step_graph_matches_eager_steps():
    """Three optimizer steps of the reference's training step (forward, QFormerLoss, backward, AdamW) replayed from
    one CUDA graph give the same loss sequence and the same weights as the same three steps enqueued eagerly."""
    from unirec_b200 import _lib, synth
    from unirec_b200.training import TrainStepGraph, qformer_loss
    B = 64
    batches = [synth.item_fields(batch=B, num_fields=6, dim=256, seed=80 + i, clip_field=2, presence=0.8) for i in range(3)]
    batches = [(x.to(DEV), m.to(DEV)) for x, m in batches]
    pos, neg = _randn(B, 256, seed=90), _randn(B, 256, seed=91)

    eager = _small_train_model(0.0)
    opt_e = torch.optim.AdamW(eager.parameters(), lr=1e-3, fused=True)
    losses_e = []
    for x, m in batches:
        loss = qformer_loss(eager(x, m), x, m, pos, neg)
        loss.backward()
        opt_e.step()
        opt_e.zero_grad(set_to_none=True)
        losses_e.append(float(loss))

    model = _small_train_model(0.0)
    opt_g = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    tg = TrainStepGraph(model, batches[0][0], batches[0][1])
    assert len(tg.grads) > 60
    losses_g = []
    for x, m in batches:
        before = _lib.launch_count()
        loss = tg.step(x, m, pos, neg)
        assert _lib.launch_count() == before          # a replay enqueues nothing through the Python wrappers
        opt_g.step()
        losses_g.append(float(loss))
    print("eager", losses_e, "graph", losses_g)
    for a, b in zip(losses_e, losses_g):
        assert abs(a - b) <= 2e-3 * abs(a) + 1e-4, (losses_e, losses_g)
    assert losses_g[0] != losses_g[1]                  # the replays really consumed new inputs / new weights
    # Adam's first updates are +-lr whatever the gradient's size, so an element whose gradient is rounding noise
    # (split-K atomics reorder fp32 sums) may move the other way: bound the worst case by 3 steps x 2 lr and require
    # that such elements are rare
    worst, total, n_el = 0.0, 0.0, 0
    for (n, pe), (_, pg) in zip(eager.named_parameters(), model.named_parameters()):
        d = (pe - pg).abs()
        worst, total, n_el = max(worst, float(d.max())), total + float(d.sum()), n_el + d.numel()
    assert worst <= 6.5e-3 and total / n_el < 1e-4, (worst, total / n_el)


def test_train_step_graph_draws_fresh_dropout_masks_per_replay():
    """With dropout 0.2 the seed frozen into the captured launches is offset by a device counter that the graph
    advances: replays on the same batch and weights give different losses, and resetting the counter reproduces one."""
    from unirec_b200 import synth
    from unirec_b200.training import TrainStepGraph
    B = 64
    x, m = synth.item_fields(batch=B, num_fields=6, dim=256, seed=85, clip_field=2, presence=0.8)
    x, m = x.to(DEV), m.to(DEV)
    model = _small_train_model(0.2)
    model.dropout_seed = 999
    tg = TrainStepGraph(model, x, m, faithful=True)
    tg.seed_offset.zero_()
    a = float(tg.step(x, m, x, x, m, m))
    grad_a = tg.grad_tensors()[5].clone()
    b = float(tg.step(x, m, x, x, m, m))
    assert int(tg.seed_offset) == 2 * TrainStepGraph.SEED_STRIDE
    assert abs(a - b) > 1e-4 * abs(a), (a, b)          # different masks
    tg.seed_offset.zero_()
    c = float(tg.step(x, m, x, x, m, m))
    assert abs(a - c) <= 1e-5 * abs(a) + 1e-6, (a, c)  # same masks (split-K atomics reorder fp32 sums only)
    torch.testing.assert_close(tg.grad_tensors()[5], grad_a, rtol=1e-3, atol=1e-5)


def test_train_step_graph_faithful_uses_each_items_own_mask():
    """The reference's step masks the positive and the negative item with THEIR attention masks
    (training/item_qformer_training.py:123-124).  Dropout 0: the faithful graph step equals the eager step that runs
    the three forwards with three different masks, and differs from the one that reuses the anchor's mask."""
    from unirec_b200 import synth
    from unirec_b200.training import TrainStepGraph, qformer_loss
    B = 64
    xs, ms = [], []
    for sd_ in (91, 92, 93):
        x_, m_ = synth.item_fields(batch=B, num_fields=6, dim=256, seed=sd_, clip_field=2, presence=0.6)
        xs.append(x_.to(DEV))
        ms.append(m_.to(DEV))
    model = _small_train_model(0.0)

    # the triplet term alone: the reconstruction term (two orders of magnitude larger) does not depend on these masks
    lk = dict(recon_weight=0.0, contrastive_weight=1.0)

    def eager_loss(mp, mn):
        out = model(xs[0], ms[0])
        with torch.no_grad():
            p = model(xs[1], mp)["item_representation"]
            n = model(xs[2], mn)["item_representation"]
        return float(qformer_loss(out, xs[0], ms[0], p, n, **lk))

    own, anchors = eager_loss(ms[1], ms[2]), eager_loss(ms[0], ms[0])
    gap = abs(own - anchors)
    assert gap > 1e-3 * abs(own), (own, anchors)           # the masks matter on this input
    import gc
    gc.collect()                                           # no eager autograd graph may be alive at capture time
    tg = TrainStepGraph(model, xs[0], ms[0], faithful=True, loss_kwargs=lk)
    got = float(tg.step(xs[0], ms[0], xs[1], xs[2], ms[1], ms[2]))
    assert abs(got - own) <= 0.05 * gap, (got, own, anchors)
    with pytest.raises(ValueError):
        tg.step(xs[0], ms[0], xs[1], xs[2])                # masks of the positive / negative items are required


@pytest.mark.parametrize("H", [256, 1024])
def test_layernorm_backward_fused_dropout_and_bias_sums(H):
    """unirec_layernorm_backward_fused: dx is the plain kernel's dx, dx_drop is dropout_backward(dx) bit for bit, dbias
    is the column sum of what the consumer GEMMs read (dx_drop, or dx without dropout)."""
    from unirec_b200 import ops
    rows = 2500
    x = _randn(rows, H, seed=41, scale=2.0, dtype=torch.bfloat16)
    dy = _randn(rows, H, seed=42, dtype=torch.bfloat16)
    dy2 = _randn(rows, H, seed=43, dtype=torch.bfloat16)
    g = _randn(H, seed=44, scale=0.1) + 1.0
    st = (ops.dropout_threshold(0.2), 777, 21)

    def zeros():
        return torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    dg0, db0 = zeros()
    dx0 = ops.layernorm_backward(x, dy, g, 1e-12, dg0, db0, dy2=dy2)
    dg1, db1 = zeros()
    dbias1 = torch.zeros(H, device=DEV)
    dx1, dxd1 = ops.layernorm_backward(x, dy, g, 1e-12, dg1, db1, dy2=dy2, dropout=st, dbias=dbias1)
    assert torch.equal(dx1, dx0)
    assert torch.equal(dxd1, ops.dropout_backward(dx0, st))
    torch.testing.assert_close(dg1, dg0, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(db1, db0, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(dbias1, dxd1.float().sum(0), rtol=1e-4, atol=2e-3)
    dg2, db2 = zeros()
    dbias2 = torch.full((H,), 0.5, device=DEV)
    dx2 = ops.layernorm_backward(x, dy, g, 1e-12, dg2, db2, dy2=dy2, dbias=dbias2)       # no dropout: sums of dx, accumulated
    assert torch.equal(dx2, dx0)
    torch.testing.assert_close(dbias2, 0.5 + dx0.float().sum(0), rtol=1e-4, atol=2e-3)
