"""GPU parity of the joint-trainer hooks (unirec_b200/joint.py, csrc/list_scoring.cu) against the golden vectors of
the unmodified reference classes and against the CPU oracle on seeded inputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "joint_scoring.npz")


def _golden():
    z = np.load(GOLDEN)
    return z, [torch.from_numpy(z[k]) for k in ("users", "pos", "negs", "masks", "lens")]


@pytest.mark.parametrize("T", [0.07, 1.0])
def test_infonce_matches_reference_golden(T):
    from unirec_b200.joint import InfoNCELoss
    z, (users, pos, negs, masks, lens) = _golden()
    crit = InfoNCELoss(temperature=T)
    u, p, n, m = users.to(DEV), pos.to(DEV), negs.to(DEV), masks.to(DEV)
    per = crit.per_user(u, p, n, m).cpu()
    ref = torch.from_numpy(z[f"loss_per_user_T{T}"])
    assert float((per - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())), (per, ref)
    assert abs(float(crit(u, p, n, m)) - float(z[f"loss_masked_T{T}"])) <= 1e-4 * max(1.0, abs(float(z[f"loss_masked_T{T}"])))
    assert abs(float(crit(u, p, n)) - float(z[f"loss_full_T{T}"])) <= 1e-4 * max(1.0, abs(float(z[f"loss_full_T{T}"])))


def test_batch_mrr_matches_reference_golden():
    from unirec_b200.joint import batch_mrr
    z, (users, pos, negs, masks, lens) = _golden()
    neg_list = [negs[i, :int(lens[i])] for i in range(len(users))]          # CPU tensors, like the validation collate
    got = batch_mrr(users.to(DEV), pos.to(DEV), neg_list)
    assert got == pytest.approx(z["mrr"].tolist(), abs=1e-7)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,C,D", [(300, 100, 1024), (7, 33, 256), (64, 1, 4096), (5, 31, 8)])
def test_list_scores_padded_and_ragged_match_oracle(B, C, D, dtype):
    """Similarities, InfoNCE loss and ranks on seeded inputs: padded lists with and without a mask, and the same lists
    in ragged (CSR) form give the same numbers, all equal to the fp32 oracle on the same (rounded) inputs."""
    from oracle import joint_oracle as JO
    from unirec_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + C)
    users = torch.randn(B, D, generator=g).to(dtype)
    pos = (users.float() * 0.3 + torch.randn(B, D, generator=g)).to(dtype)
    negs = torch.randn(B, C, D, generator=g).to(dtype)
    lens = torch.randint(1, C + 1, (B,), generator=g)
    lens[0] = C
    masks = torch.arange(C).unsqueeze(0) < lens.unsqueeze(1)
    neg_list = [negs[i, :int(lens[i])] for i in range(B)]
    ref_sims = JO.list_similarities(users, pos, neg_list)
    tol = 3e-5 if dtype == torch.float32 else 3e-5      # fp32 accumulation either way; inputs identical
    u, p, n, m = users.to(DEV), pos.to(DEV), negs.to(DEV), masks.to(DEV)
    sims_m, inv = ops.list_scores(u, p, n, mask=m)
    offs = torch.zeros(B + 1, dtype=torch.int64)
    offs[1:] = lens.cumsum(0)
    sims_r, _ = ops.list_scores(u, p, torch.cat(neg_list, 0).to(DEV), offsets=offs.to(DEV), max_list=C)
    assert torch.equal(sims_m, sims_r)
    sims_cpu = sims_m.cpu()
    for i in range(B):
        k = int(lens[i]) + 1
        assert float((sims_cpu[i, :k] - ref_sims[i]).abs().max()) <= tol
        assert bool(torch.isinf(sims_cpu[i, k:]).all()) and bool((sims_cpu[i, k:] < 0).all())
    norms = torch.cat([pos.float().unsqueeze(1), negs.float()], 1).norm(dim=-1)
    got_inv = inv.cpu()
    valid = torch.cat([torch.ones(B, 1, dtype=torch.bool), masks], 1)
    torch.testing.assert_close(got_inv[valid], (1.0 / norms)[valid], rtol=1e-5, atol=0)
    assert float(got_inv[~valid].abs().sum()) == 0.0
    for T in (0.07, 0.5):
        loss, rank = ops.infonce_rank(sims_m, T)
        ref_loss = JO.infonce_per_user(users, pos, negs, masks, T)
        torch.testing.assert_close(loss.cpu(), ref_loss, rtol=2e-4, atol=2e-3 if T < 0.1 else 3e-4)
    ref_rank = torch.tensor([1 + int((s[1:] > s[0]).sum()) for s in ref_sims], dtype=torch.int32)
    # a rank may differ only where a negative ties with the positive within the similarity tolerance
    bad = [i for i in range(B) if int(rank[i]) != int(ref_rank[i])
           and float((ref_sims[i][1:] - ref_sims[i][0]).abs().min()) > 2 * tol]
    assert not bad, bad
    # no mask = every padded slot is a real candidate
    sims_f, _ = ops.list_scores(u, p, n)
    full = JO.list_similarities(users, pos, [negs[i] for i in range(B)])
    assert float((sims_f.cpu() - torch.stack(full)).abs().max()) <= tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_infonce_backward_matches_oracle_autograd(dtype):
    from oracle import joint_oracle as JO
    from unirec_b200.joint import InfoNCELoss
    B, C, D, T = 37, 50, 512, 0.07
    g = torch.Generator().manual_seed(11)
    users = torch.randn(B, D, generator=g).to(dtype)
    pos = (users.float() * 0.3 + torch.randn(B, D, generator=g)).to(dtype)
    negs = torch.randn(B, C, D, generator=g).to(dtype)
    masks = torch.rand(B, C, generator=g) < 0.8
    masks[:, 0] = True
    ru, rp, rn = (t.float().clone().requires_grad_() for t in (users, pos, negs))
    JO.infonce_loss(ru, rp, rn, masks, T).backward()
    u, p, n = (t.to(DEV).requires_grad_() for t in (users, pos, negs))
    loss = InfoNCELoss(T)(u, p, n, masks.to(DEV))
    loss.backward()
    rt = 1e-3 if dtype == torch.float32 else 2e-2       # bf16: the returned gradients are rounded to bf16
    for got, ref, name in ((u.grad, ru.grad, "users"), (p.grad, rp.grad, "pos"), (n.grad, rn.grad, "negs")):
        assert got.dtype == dtype
        err = float((got.float().cpu() - ref).abs().max())
        assert err <= rt * float(ref.abs().max()) + 1e-7, (name, err, float(ref.abs().max()))
    assert float(n.grad.float().cpu()[~masks].abs().sum()) == 0.0       # padding receives no gradient
    # only the user side needs a gradient in the reference (item embeddings are data): no list gradient is built
    u2 = users.to(DEV).requires_grad_()
    InfoNCELoss(T)(u2, pos.to(DEV), negs.to(DEV), masks.to(DEV)).backward()
    assert float((u2.grad.float() - u.grad.float()).abs().max()) <= 1e-6 + (0 if dtype == torch.float32 else 1e-3)


@pytest.mark.parametrize("src,dst", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                     (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32)])
def test_inject_history_tokens_bit_exact(src, dst):
    from oracle import joint_oracle as JO
    from unirec_b200.joint import inject_history_tokens
    B, S, nh, Q, Hd = 5, 700, 10, 32, 1024
    g = torch.Generator().manual_seed(3)
    token_ids = (151_700 + torch.randperm(nh * Q, generator=g)).view(nh, Q)       # not sorted, not contiguous in (i, j)
    input_ids = torch.randint(0, 151_000, (B, S), generator=g)
    for b in range(B):            # each placeholder once at a random position, a few twice, some users miss some
        perm = torch.randperm(S, generator=g)[: nh * Q + 5]
        input_ids[b, perm[: nh * Q]] = token_ids.reshape(-1)
        input_ids[b, perm[nh * Q:]] = token_ids.reshape(-1)[:5]
        if b % 2:
            input_ids[b, perm[7:19]] = 42
    text = torch.randn(B, S, Hd, generator=g).to(dst)
    toks = torch.randn(B, nh, Q, Hd, generator=g).to(src)
    ref = JO.inject_tokens(text, input_ids, token_ids, toks)
    got = inject_history_tokens(text.to(DEV), input_ids.to(DEV), token_ids, toks.to(DEV))
    assert torch.equal(got.cpu(), ref)


def test_inject_history_tokens_matches_reference_forward_golden():
    """Golden `inj_*` = what the unmodified reference forward handed to the LLM (stub tokenizer / LLM, see
    oracle/pin_joint_against_reference.py::pin_injection): the kernel must reproduce it bit for bit."""
    from unirec_b200.joint import inject_history_tokens
    z = np.load(GOLDEN)
    got = inject_history_tokens(torch.from_numpy(z["inj_text"]).to(DEV), torch.from_numpy(z["inj_input_ids"]).to(DEV),
                                torch.from_numpy(z["inj_token_ids"]), torch.from_numpy(z["inj_tokens"]).to(DEV))
    assert torch.equal(got.cpu(), torch.from_numpy(z["inj_out"]))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_inject_history_tokens_gradients_match_index_assign(dtype):
    """The overwrite is differentiable in the reference (:160-171, `text_embeds[b, positions] = query_embeddings[b]`):
    gradients of a downstream loss must reach the query tokens (the only path to the item Q-Former) and vanish at the
    overwritten rows of the text embeddings.  Compared with torch autograd through the oracle's loop, exactly (fp32)."""
    from oracle import joint_oracle as JO
    from unirec_b200.joint import inject_history_tokens
    B, S, nh, Q, Hd = 3, 200, 4, 8, 64
    g = torch.Generator().manual_seed(11)
    token_ids = (151_700 + torch.randperm(nh * Q, generator=g)).view(nh, Q)
    input_ids = torch.randint(0, 151_000, (B, S), generator=g)
    for b in range(B):
        perm = torch.randperm(S, generator=g)[: nh * Q + 3]
        input_ids[b, perm[: nh * Q]] = token_ids.reshape(-1)
        input_ids[b, perm[nh * Q:]] = token_ids.reshape(-1)[:3]          # three placeholders appear twice
        if b == 1:
            input_ids[b, perm[2:6]] = 7                                  # this user misses four placeholders
    base = torch.randn(B, S, Hd, generator=g).to(dtype)
    toks = torch.randn(B, nh, Q, Hd, generator=g).to(dtype)
    w = torch.randn(B, S, Hd, generator=g)

    def run(dev, inject):
        e = base.to(dev).clone().requires_grad_(True)
        t = toks.to(dev).clone().requires_grad_(True)
        out = inject(e * 1.0, input_ids.to(dev), token_ids, t)           # e * 1.0: a non-leaf, like an embedding lookup
        (out.float() * w.to(dev)).sum().backward()
        return out.detach().cpu(), e.grad.cpu(), t.grad.cpu()

    ref_out, ref_de, ref_dt = run("cpu", JO.inject_tokens)
    out, de, dt = run(DEV, inject_history_tokens)
    assert torch.equal(out, ref_out)
    tol = 0 if dtype == torch.float32 else 2e-2
    assert float((de.float() - ref_de.float()).abs().max()) <= tol
    assert float((dt.float() - ref_dt.float()).abs().max()) <= tol * 4 + 1e-6    # duplicates: two fp32 adds, order-free
    assert float(dt.abs().sum()) > 0            # the Q-Former side does get a gradient


def test_history_query_tokens_is_the_item_qformer_on_flattened_history():
    from unirec_b200 import synth
    from unirec_b200.joint import history_query_tokens
    from unirec_b200.modules import QFormerForItemRepresentation
    model = QFormerForItemRepresentation(hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                                         intermediate_size=512, field_embedding_dim=256, num_fields=6).to(DEV).eval()
    x, m = synth.item_fields(batch=3 * 4, num_fields=6, dim=256, seed=5, clip_field=2, presence=0.8)
    x, m = x.to(DEV), m.to(DEV)
    tok = history_query_tokens(model, x.view(3, 4, 6, 256), m.view(3, 4, 6))
    assert tuple(tok.shape) == (3, 4, 32, 256)
    assert torch.equal(tok.view(12, 32, 256), model(x, m)["query_outputs"])


def test_list_scores_rejects_bad_arguments():
    from unirec_b200 import ops
    u = torch.randn(4, 64, device=DEV)
    with pytest.raises(RuntimeError):
        ops.list_scores(u.cpu(), u.cpu(), torch.randn(4, 3, 64))                      # no CPU path
    with pytest.raises(RuntimeError):
        ops.list_scores(u, u, torch.randn(4, 3, 64, device=DEV).to(torch.bfloat16))   # mixed dtypes
    with pytest.raises(RuntimeError):
        ops.list_scores(u[:, :60], u[:, :60], torch.randn(4, 3, 60, device=DEV))      # D % 8 != 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_reconstruction_metrics_kernel_matches_oracle(dtype):
    from oracle import eval_oracle as EO
    from unirec_b200 import ops
    g = torch.Generator().manual_seed(21)
    B, F_, E = 300, 14, 1024
    orig = torch.randn(B, F_, E, generator=g)
    orig[:, 7, 768:] = 0
    rec = (orig * 0.8 + 0.5 * torch.randn(B, F_, E, generator=g)).to(dtype)
    mask = (torch.rand(B, F_, generator=g) < 0.7).long()
    mask[3] = 0
    loss, cos_sum, n = EO.batch_metrics(rec, orig, mask)
    acc = ops.reconstruction_metrics(rec.to(DEV), orig.to(DEV), mask.to(DEV)).cpu()
    assert int(acc[2]) == n == int(mask.sum())
    assert abs(float(acc[0]) / n - loss) <= 1e-5 * loss
    assert abs(float(acc[1]) - cos_sum) <= 1e-5 * abs(cos_sum)
    # accumulation over batches into the same three doubles
    acc2 = torch.zeros(3, device=DEV, dtype=torch.float64)
    for lo in range(0, B, 128):
        ops.reconstruction_metrics(rec[lo:lo + 128].to(DEV), orig[lo:lo + 128].to(DEV), mask[lo:lo + 128].to(DEV), acc2)
    torch.testing.assert_close(acc2.cpu(), acc, rtol=1e-6, atol=0)


def test_evaluate_reconstruction_matches_reference_golden():
    """The evaluation loop on the CUDA path against the two numbers of the unmodified reference function."""
    from tests.golden_cases import EVAL_CASE, ITEM_CASES
    from unirec_b200 import synth
    from unirec_b200.evaluation import evaluate_reconstruction
    from unirec_b200.modules import QFormerForItemRepresentation
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_metrics.npz"))
    c = ITEM_CASES[EVAL_CASE["item_case"]]
    mk = c["model"]
    sd = synth.item_qformer_state_dict(**mk, seed=c["seed"], attn_std=c["attn_std"])
    model = QFormerForItemRepresentation(hidden_size=mk["hidden"], num_hidden_layers=mk["layers"],
                                         num_attention_heads=c["heads"], intermediate_size=mk["inter"],
                                         num_query_tokens=mk["num_query"], field_embedding_dim=mk["field_dim"],
                                         num_fields=mk["num_fields"])
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    x, mask = synth.item_fields(**EVAL_CASE["input"])                       # host tensors, like the cached files
    got = evaluate_reconstruction(model, x, mask, batch_size=EVAL_CASE["batch_size"])
    # bf16 activations: the reconstruction differs from fp32 by ~1e-2 relative per element; the loss is a mean of
    # squares dominated by the targets' own energy
    assert abs(got["val_recon_loss"] - float(z["val_recon_loss"])) <= 5e-3 * float(z["val_recon_loss"]), got
    assert abs(got["avg_cosine_similarity"] - float(z["avg_cosine_similarity"])) <= 2e-3, got


def test_list_scores_more_users_than_one_grid():
    """B > 65535 users (the grid's y limit): the wrapper slices the batch; padded and ragged forms agree with torch."""
    from unirec_b200 import ops
    B, C, D = 70_000, 5, 64
    g = torch.Generator(device=DEV).manual_seed(2)
    u = torch.randn(B, D, device=DEV, generator=g)
    p = torch.randn(B, D, device=DEV, generator=g)
    n = torch.randn(B, C, D, device=DEV, generator=g)
    sims, _ = ops.list_scores(u, p, n)
    ref = torch.einsum("bd,bcd->bc", torch.nn.functional.normalize(u, dim=-1),
                       torch.nn.functional.normalize(torch.cat([p.unsqueeze(1), n], 1), dim=-1))
    assert float((sims - ref).abs().max()) <= 3e-5
    offs = torch.arange(B + 1, device=DEV, dtype=torch.int64) * C
    sims_r, _ = ops.list_scores(u, p, n.view(B * C, D), offsets=offs, max_list=C)
    assert torch.equal(sims_r, sims)
