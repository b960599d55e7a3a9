"""GPU parity of fused cosine scoring + top-k, the top-k merge and the nested ranker against the oracle.

Scores: the kernel multiplies bf16 inputs exactly and accumulates in fp32, the oracle does the same in
fp32 on the same bf16-rounded inputs, so scores agree to fp32 summation-order error (atol 2e-5).  Indices
must be identical except where the oracle's neighbouring scores tie within that tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import qformer_oracle as O
from tests.golden_cases import SCORING_CASE
from unirec_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-5


def _assert_topk_matches(s, i, users_bf, cands_bf, k):
    full = O.cosine_scores(users_bf.cpu().float(), cands_bf.cpu().float())     # [B, N] oracle
    ks = min(k, full.shape[1])
    ref_s, ref_i = torch.topk(full, ks, dim=-1)
    s, i = s.cpu(), i.cpu()
    assert torch.allclose(s[:, :ks], ref_s, atol=TOL, rtol=0), float((s[:, :ks] - ref_s).abs().max())
    # every returned index really has the returned score, no duplicates, descending order
    got = torch.gather(full, 1, i[:, :ks])
    assert torch.allclose(got, s[:, :ks], atol=TOL, rtol=0)
    assert all(len(set(r.tolist())) == ks for r in i[:, :ks])
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    # index identity wherever the oracle's gap to both neighbours exceeds the tolerance
    gap_ok = torch.ones_like(ref_s, dtype=torch.bool)
    d = (ref_s[:, :-1] - ref_s[:, 1:]) > 4 * TOL
    gap_ok[:, 1:] &= d
    gap_ok[:, :-1] &= d
    kth_gap = (ref_s[:, -1] - torch.topk(full, min(ks + 1, full.shape[1]), dim=-1)[0][:, -1]) > 4 * TOL
    gap_ok[:, -1] &= kth_gap if ks < full.shape[1] else True
    assert bool((i[:, :ks][gap_ok] == ref_i[gap_ok]).all())
    if ks < k:
        assert bool((i[:, ks:] == -1).all()) and bool(torch.isinf(s[:, ks:]).all())


def test_scoring_golden_case():
    from unirec_b200 import ops
    c = SCORING_CASE
    u = synth.normal("score_users", (c["users"], c["dim"]), c["seed"]).to(torch.bfloat16).to(DEV)
    C = synth.normal("score_cands", (c["cands"], c["dim"]), c["seed"]).to(torch.bfloat16).to(DEV)
    s, i = ops.score_topk(u, C, c["k"])
    _assert_topk_matches(s, i, u, C, c["k"])
    # against the reference-produced golden (fp32 inputs): bf16 input rounding moves cosine by <~ 4e-3
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "scoring.npz"))
    assert np.abs(s.cpu().numpy() - g["scores"]).max() < 4e-3
    overlap = np.mean([len(set(a.tolist()) & set(b.tolist())) / c["k"] for a, b in zip(i.cpu().numpy(), g["indices"])])
    assert overlap > 0.9


@pytest.mark.parametrize("B,N,k", [(1, 300, 100), (130, 5000, 100), (64, 70001, 100), (300, 40960, 128), (5, 50, 100),
                                   (17, 4096, 1), (3, 32768, 100), (2, 32769, 7), (40, 150001, 100)])
def test_score_topk_random(B, N, k):
    from unirec_b200 import ops
    g = torch.Generator().manual_seed(B * 31 + N)
    u = torch.randn(B, 256, generator=g).to(torch.bfloat16).to(DEV)
    C = (torch.randn(N, 256, generator=g) * torch.rand(N, 1, generator=g) * 3).to(torch.bfloat16).to(DEV)
    s, i = ops.score_topk(u, C, k, index_base=0)
    _assert_topk_matches(s, i, u, C, k)
    s2, i2 = ops.score_topk(u, C, k, index_base=1000)
    assert torch.equal(i2[i >= 0], i[i >= 0] + 1000)


def test_score_topk_adversarial_order_and_ties():
    """Candidates sorted by ascending similarity to user 0 (every new tile beats the running threshold,
    forcing repeated list compactions), plus exact duplicates (ties)."""
    from unirec_b200 import ops
    g = torch.Generator().manual_seed(7)
    N, D = 60000, 128
    u = torch.randn(4, D, generator=g).to(torch.bfloat16)
    C = torch.randn(N, D, generator=g).to(torch.bfloat16)
    order = torch.argsort(O.cosine_scores(u[:1].float(), C.float())[0])
    C = C[order].contiguous()
    C[100:400] = C[N - 1]       # 300 copies of the best candidate for user 0: massive tie at the top
    s, i = ops.score_topk(u.to(DEV), C.to(DEV), 100)
    _assert_topk_matches(s, i, u, C, 100)
    assert float(s[0, 0]) == pytest.approx(float(s[0, 99]), abs=TOL)   # all 100 winners are the tied copies


def test_score_topk_adversarial_sample():
    """Large pool (sampled start threshold).  The sampled candidate tiles (every 17th tile of 256 rows for
    N = 140000) hold only candidates that are bad for user 0, every other tile is good: the sampled threshold
    is useless for that user, ~all candidates pass the filter and the exact list-compaction path must keep
    the result right.  User 1 sees the opposite (sample = the best candidates: threshold is the true one)."""
    from unirec_b200 import ops
    g = torch.Generator().manual_seed(21)
    N, D, k = 140000, 64, 100
    u = torch.randn(3, D, generator=g)
    C = torch.randn(N, D, generator=g)
    tile = torch.arange(N) // 256
    sampled = (tile % 17 == 0) & (tile // 17 < 32)
    C[sampled] = -u[0] + 0.3 * C[sampled]
    C[~sampled] = 0.5 * u[0] + C[~sampled]
    u, C = u.to(torch.bfloat16), C.to(torch.bfloat16)
    s, i = ops.score_topk(u.to(DEV), C.to(DEV), k)
    _assert_topk_matches(s, i, u, C, k)


def test_score_topk_full_size_properties():
    """BASELINE size (1 M candidates x 1024, top-100): properties that do not need the CPU oracle's full [B, N]
    matrix.  (1) lists are descending with unique in-range indices; (2) every returned score is the exact cosine of its
    (user, candidate) pair; (3) completeness - no candidate outside a list beats the list's k-th score by more than
    the tolerance (checked by a chunked fp32 torch.matmul sweep, used as a checker only); (4) scoring the pool as 8 row
    shards + unirec_topk_merge (the multi-GPU path, config 5) gives the same lists; (5) the same call twice is
    bit-identical."""
    from unirec_b200 import ops
    N, D, B, k, G = 1_000_000, 1024, 192, 100, 8
    gen = torch.Generator(device=DEV).manual_seed(77)
    C = torch.randn(N, D, device=DEV, generator=gen).to(torch.bfloat16)
    u = torch.randn(B, D, device=DEV, generator=gen).to(torch.bfloat16)
    s, i = ops.score_topk(u, C, k)
    s2, i2 = ops.score_topk(u, C, k)
    assert torch.equal(s, s2) and torch.equal(i, i2)                                      # (5)
    assert bool((s[:, :-1] >= s[:, 1:]).all()) and int(i.min()) >= 0 and int(i.max()) < N   # (1)
    assert all(len(set(r.tolist())) == k for r in i.cpu())
    un = torch.nn.functional.normalize(u.float(), dim=-1)
    picked = torch.nn.functional.normalize(C[i.reshape(-1)].float(), dim=-1).view(B, k, D)
    exact = torch.einsum("bd,bkd->bk", un, picked)
    assert torch.allclose(exact, s, atol=TOL, rtol=0), float((exact - s).abs().max())      # (2)
    kth = s[:, -1:]
    beating = torch.zeros(B, device=DEV, dtype=torch.long)
    for c0 in range(0, N, 65536):                                                           # (3)
        cn = torch.nn.functional.normalize(C[c0:c0 + 65536].float(), dim=-1)
        beating += ((un @ cn.t()) > kth + 4 * TOL).sum(dim=1)
    assert int(beating.max()) <= k - 1, int(beating.max())
    parts_s, parts_i = [], []
    for r in range(G):                                                                      # (4)
        lo, hi = r * N // G, (r + 1) * N // G
        ps, pi = ops.score_topk(u, C[lo:hi], k, index_base=lo)
        parts_s.append(ps)
        parts_i.append(pi)
    ms, mi = ops.topk_merge(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.allclose(ms, s, atol=TOL, rtol=0)
    same = (mi == i)
    gap = torch.ones_like(same)
    d = (s[:, :-1] - s[:, 1:]) > 4 * TOL
    gap[:, 1:] &= d
    gap[:, :-1] &= d
    assert bool(same[gap].all())


def test_topk_merge_matches_global_topk():
    from unirec_b200 import ops
    g = torch.Generator().manual_seed(11)
    B, N, k, G = 33, 8000, 100, 4
    u = torch.randn(B, 128, generator=g).to(torch.bfloat16).to(DEV)
    C = torch.randn(N, 128, generator=g).to(torch.bfloat16).to(DEV)
    parts_s, parts_i = [], []
    for r in range(G):
        lo, hi = r * N // G, (r + 1) * N // G
        s, i = ops.score_topk(u, C[lo:hi], k, index_base=lo)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = ops.topk_merge(torch.stack(parts_s), torch.stack(parts_i))
    _assert_topk_matches(ms, mi, u, C, k)


def test_nested_ranker_end_to_end_small():
    """item tokens -> user sequence -> UserQFormer -> pooled vector -> top-k, vs the oracle chain."""
    from unirec_b200 import ops
    from unirec_b200.modules import UserQFormer
    from unirec_b200.pipeline import NestedRanker
    mk = dict(hidden=256, layers=2, inter=512, num_query=64, input_dim=256, num_predict=8)
    sd = synth.user_qformer_state_dict(**mk, seed=31, attn_std=0.1)
    um = UserQFormer(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                     num_query_tokens=64, input_embedding_dim=256, num_item_tokens_to_predict=8)
    um.load_state_dict(sd, strict=True)
    um = um.to(DEV).eval()
    N, Q, D, B, Hmax, k = 3000, 32, 256, 9, 6, 100
    table = synth.normal("tok_table", (N, Q, D), 32).to(torch.bfloat16)
    cands = table.float().mean(dim=1).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(33)
    history = torch.randint(0, N, (B, Hmax), generator=gen)
    lengths = torch.randint(1, Hmax + 1, (B,), generator=gen).to(torch.int32)
    ranker = NestedRanker(um, table.to(DEV), cands.to(DEV), k=k)
    uvec = ranker.encode_users(history.to(DEV), lengths.to(DEV))
    # oracle chain on CPU (fp32) from the same bf16 table
    seq, mask = O.build_user_sequences(table.float(), history, lengths.long())
    pred = O.user_qformer_forward(sd, seq, mask, num_heads=4, num_item_tokens_to_predict=8)
    u_ref = O.pooled_scoring_vector(pred)
    cos = torch.nn.functional.cosine_similarity(uvec.float().cpu(), u_ref, dim=-1)
    print("user-vector cosine vs oracle:", cos.min().item())
    assert float(cos.min()) > 0.999
    s, i = ranker.rank(uvec)
    _assert_topk_matches(s, i, uvec, cands, k)           # exact w.r.t. the vectors actually scored
    ref_s, ref_i = O.cosine_topk(u_ref, cands.float(), k)
    overlap = np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(i.cpu().numpy(), ref_i.numpy())])
    print("top-100 overlap with the fp32 oracle chain:", overlap)
    assert overlap > 0.85 and float((s.cpu() - ref_s).abs().max()) < 2e-2


@pytest.mark.parametrize("B,N", [(32768, 125_000), (16384, 250_000)])
def test_score_topk_multi_gpu_rank_shapes(B, N):
    """The per-rank scoring call of the 8- and 4-GPU runs of config 5: ALL users of the group (4096 per GPU) against this
    rank's 1/G of the 1 M-row pool (B x N > 2^32 score pairs).  Properties as in the full-size test, checked on a sample
    of user rows spread over the whole batch: descending unique in-range lists, exact cosines, completeness."""
    from unirec_b200 import ops
    D, k = 1024, 100
    gen = torch.Generator(device=DEV).manual_seed(B + N)
    C = torch.randn(N, D, device=DEV, generator=gen).to(torch.bfloat16)
    u = torch.randn(B, D, device=DEV, generator=gen).to(torch.bfloat16)
    s, i = ops.score_topk(u, C, k, index_base=7 * N)
    i = i - 7 * N
    assert tuple(s.shape) == (B, k) and bool((s[:, :-1] >= s[:, 1:]).all())
    assert int(i.min()) >= 0 and int(i.max()) < N
    rows = torch.cat([torch.arange(0, 40), torch.arange(B // 2 - 20, B // 2 + 20), torch.arange(B - 40, B),
                      torch.randint(0, B, (136,), generator=torch.Generator().manual_seed(1))]).to(DEV)
    us, ss, ii = u[rows], s[rows], i[rows]
    assert all(len(set(r.tolist())) == k for r in ii.cpu())
    un = torch.nn.functional.normalize(us.float(), dim=-1)
    picked = torch.nn.functional.normalize(C[ii.reshape(-1)].float(), dim=-1).view(len(rows), k, D)
    exact = torch.einsum("bd,bkd->bk", un, picked)
    assert torch.allclose(exact, ss, atol=TOL, rtol=0), float((exact - ss).abs().max())
    beating = torch.zeros(len(rows), device=DEV, dtype=torch.long)
    for c0 in range(0, N, 65536):
        cn = torch.nn.functional.normalize(C[c0:c0 + 65536].float(), dim=-1)
        beating += ((un @ cn.t()) > ss[:, -1:] + 4 * TOL).sum(dim=1)
    assert int(beating.max()) <= k - 1, int(beating.max())


def test_linear_gather_matches_materialised_projection():
    """The K/V projection with its A operand gathered from the item-token table: ragged histories with a
    zero-length user, garbage ids in the padding slots, user boundaries that fall inside 128-row tiles, a row count that
    is not a multiple of the 256-row tile - against fp32 torch on the same bf16 operands."""
    from unirec_b200 import ops
    N_items, K, Hmax, B, N = 500, 256, 6, 9, 1024
    S = Hmax * 32
    g = torch.Generator().manual_seed(55)
    table = torch.randn(N_items, 32, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    pad = torch.randn(S, K, generator=g).to(torch.bfloat16)
    posb = torch.randn(S, N, generator=g).to(torch.bfloat16)
    posb_ext = torch.cat([posb, posb[:128]], 0).contiguous()
    hist = torch.randint(0, N_items, (B, Hmax), generator=g)
    lengths = torch.tensor([6, 0, 3, 1, 6, 2, 5, 6, 4], dtype=torch.int32)
    for b in range(B):                                   # ids in padding slots are never read
        hist[b, int(lengths[b]):] = torch.tensor([-1, 10 ** 12, N_items, -7, 3, 0])[: Hmax - int(lengths[b])]
    A = torch.empty(B, Hmax, 32, K)
    for b in range(B):
        for j in range(Hmax):
            A[b, j] = table[hist[b, j]].float() if j < int(lengths[b]) else pad[j * 32:(j + 1) * 32].float()
    ref = A.view(B * S, K) @ W.float().t() + bias + posb.float().repeat(B, 1)
    out = ops.linear_gather(table.to(DEV), hist.to(DEV), lengths.to(DEV), pad.to(DEV), W.to(DEV), bias.to(DEV),
                            posb_ext.to(DEV))
    assert tuple(out.shape) == (B * S, N)
    torch.testing.assert_close(out.float().cpu(), ref, rtol=1e-2, atol=3e-2)
    # a valid slot whose id lies outside the table reads zero rows (TMA out-of-bounds fill), like an empty item
    hist2 = hist.clone()
    hist2[0, 2] = N_items + 5
    out2 = ops.linear_gather(table.to(DEV), hist2.to(DEV), lengths.to(DEV), pad.to(DEV), W.to(DEV), bias.to(DEV),
                             posb_ext.to(DEV))
    rows = slice(2 * 32, 3 * 32)
    torch.testing.assert_close(out2[rows].float().cpu(), (bias + posb[rows].float()), rtol=1e-2, atol=3e-2)
    assert torch.equal(out2[3 * 32:], out[3 * 32:])


def test_nested_ranker_fused_gather_matches_builder_path_and_oracle():
    """NestedRanker(fused_gather=True): user vectors from the gathered K/V projection against the oracle chain and
    against the path that materialises the user sequence (including a user with an empty history: uniform attention
    over all-padding keys)."""
    from unirec_b200.modules import UserQFormer
    from unirec_b200.pipeline import NestedRanker
    mk = dict(hidden=256, layers=2, inter=512, num_query=64, input_dim=256, num_predict=8)
    sd = synth.user_qformer_state_dict(**mk, seed=31, attn_std=0.1)
    um = UserQFormer(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                     num_query_tokens=64, input_embedding_dim=256, num_item_tokens_to_predict=8)
    um.load_state_dict(sd, strict=True)
    um = um.to(DEV).eval()
    N, Q, D, B, Hmax, k = 3000, 32, 256, 11, 6, 50
    table = synth.normal("tok_table", (N, Q, D), 32).to(torch.bfloat16)
    cands = table.float().mean(dim=1).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(34)
    history = torch.randint(0, N, (B, Hmax), generator=gen)
    lengths = torch.randint(1, Hmax + 1, (B,), generator=gen).to(torch.int32)
    lengths[4] = 0
    lengths[7] = Hmax
    plain = NestedRanker(um, table.to(DEV), cands.to(DEV), k=k)
    fused = NestedRanker(um, table.to(DEV), cands.to(DEV), k=k, fused_gather=True)
    u_plain = plain.encode_users(history.to(DEV), lengths.to(DEV)).float().cpu()
    u_fused = fused.encode_users(history.to(DEV), lengths.to(DEV)).float().cpu()
    seq, mask = O.build_user_sequences(table.float(), history, lengths.long())
    pred = O.user_qformer_forward(sd, seq, mask, num_heads=4, num_item_tokens_to_predict=8)
    u_ref = O.pooled_scoring_vector(pred)
    cos_ref = torch.nn.functional.cosine_similarity(u_fused, u_ref, dim=-1)
    cos_plain = torch.nn.functional.cosine_similarity(u_fused, u_plain, dim=-1)
    print("fused vs oracle:", cos_ref.min().item(), " fused vs builder path:", cos_plain.min().item())
    assert float(cos_ref.min()) > 0.999 and float(cos_plain.min()) > 0.999
    s, i = fused(history.to(DEV), lengths.to(DEV))
    assert tuple(s.shape) == (B, k) and bool((s[:, :-1] >= s[:, 1:]).all())


def test_unknown_history_ids_are_zero_rows_on_both_user_paths():
    """Ids outside [0, num_items) (-1 sentinels, items missing from the token table): the sequence builder treats them
    as a ZERO token row whose slot keeps its position term and stays attended - never an out-of-bounds read - and the
    gathered K/V projection does the same, so the two user-encoding paths agree on bad ids."""
    from unirec_b200 import ops
    from unirec_b200.modules import UserQFormer
    from unirec_b200.pipeline import NestedRanker
    N, Q, D, B, Hmax = 500, 32, 256, 6, 5
    table = synth.normal("tok_table_bad", (N, Q, D), 35).to(torch.bfloat16)
    gen = torch.Generator().manual_seed(36)
    history = torch.randint(0, N, (B, Hmax), generator=gen)
    history[0, 1], history[2, 0], history[3, 4], history[5, 2] = -1, N, N + 12345, -(2 ** 40)
    lengths = torch.tensor([5, 3, 5, 5, 2, 4], dtype=torch.int32)
    seq, mask = ops.build_user_sequence(table.to(DEV), history.to(DEV), lengths.to(DEV))
    # oracle with a zero item appended for the unknown ids
    table0 = torch.cat([table.float(), torch.zeros(1, Q, D)], 0)
    known = (history >= 0) & (history < N)
    ref_seq, ref_mask = O.build_user_sequences(table0, torch.where(known, history, torch.full_like(history, N)), lengths.long())
    assert torch.equal(mask.cpu(), ref_mask.float())
    assert float((seq.float().cpu() - ref_seq).abs().max()) <= 0.04
    mk = dict(hidden=256, layers=2, inter=512, num_query=64, input_dim=256, num_predict=8)
    um = UserQFormer(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                     num_query_tokens=64, input_embedding_dim=256, num_item_tokens_to_predict=8)
    um.load_state_dict(synth.user_qformer_state_dict(**mk, seed=31, attn_std=0.1), strict=True)
    um = um.to(DEV).eval()
    cands = table.float().mean(dim=1).to(torch.bfloat16).to(DEV)
    u_plain = NestedRanker(um, table.to(DEV), cands, k=10).encode_users(history.to(DEV), lengths.to(DEV)).float()
    u_fused = NestedRanker(um, table.to(DEV), cands, k=10, fused_gather=True).encode_users(history.to(DEV), lengths.to(DEV)).float()
    assert float(torch.nn.functional.cosine_similarity(u_plain, u_fused, dim=-1).min()) > 0.999


def test_invalidate_packed_picks_up_data_writes():
    """In-place writes through `.data` do not bump the version counter the bf16 weight cache is keyed on; the explicit
    `invalidate_packed()` (also run by load_state_dict and .to()) makes the next forward use the new weights."""
    from unirec_b200.modules import UserQFormer
    mk = dict(hidden=256, layers=2, inter=512, num_query=64, input_dim=256, num_predict=8)
    um = UserQFormer(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                     num_query_tokens=64, input_embedding_dim=256, num_item_tokens_to_predict=8)
    um.load_state_dict(synth.user_qformer_state_dict(**mk, seed=31, attn_std=0.1), strict=True)
    um = um.to(DEV).eval()
    x = torch.randn(3, 64, 256, device=DEV)
    m = torch.ones(3, 64, device=DEV)
    a = um(x, m).clone()
    um.qformer.encoder.layer[1].output_query.dense.weight.data.mul_(0.5)       # no version bump
    um.invalidate_packed()
    b = um(x, m).clone()
    assert float((a - b).abs().max()) > 1e-3
    sd2 = synth.user_qformer_state_dict(**mk, seed=31, attn_std=0.1)
    um.load_state_dict(sd2, strict=True)                                       # the post-hook invalidates
    assert torch.equal(um(x, m), a)
