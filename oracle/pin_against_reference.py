"""Pin the CPU oracle against the UNMODIFIED reference and (re)generate tests/golden/*.npz.

Run in the authoring container only (needs /root/reference):

    python -m oracle.pin_against_reference            # check + write golden vectors
    python -m oracle.pin_against_reference --check    # check only

For every case: synthetic weights (unirec_b200.synth, a pure function of key/shape/seed) are
loaded into the reference nn.Module with load_state_dict(strict=True), the reference runs in
eval()/no_grad()/fp32 on CPU, the oracle runs on the same state dict, and the two must agree to
2e-5.  The REFERENCE's outputs (not the oracle's) are what is stored.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import qformer_oracle as O  # noqa: E402
from oracle import reference_shim  # noqa: E402
from unirec_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
TOL = 2e-5

# name -> (constructor kwargs, synth kwargs, input kwargs).  tests/golden_cases.py mirrors this
# table (it cannot import this module on the GPU box because of the reference dependency).
from tests.golden_cases import ITEM_CASES, USER_CASES, SCORING_CASE  # noqa: E402


def _maxdiff(a, b):
    return float((a - b).abs().max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    Item, User, PE = reference_shim.load()
    os.makedirs(GOLDEN, exist_ok=True)
    worst = 0.0

    for name, c in ITEM_CASES.items():
        sd = synth.item_qformer_state_dict(**c["model"], seed=c["seed"], attn_std=c["attn_std"])
        m = Item(hidden_size=c["model"]["hidden"], num_hidden_layers=c["model"]["layers"],
                 num_attention_heads=c["heads"], intermediate_size=c["model"]["inter"],
                 num_query_tokens=c["model"]["num_query"], field_embedding_dim=c["model"]["field_dim"],
                 num_fields=c["model"]["num_fields"]).eval()
        m.load_state_dict(sd, strict=True)
        x, mask = synth.item_fields(**c["input"])
        with torch.no_grad():
            ref = m(x, mask)
            ref_nomask = m(x[:1], None)
            ora = O.item_qformer_forward(sd, x, mask, num_heads=c["heads"])
        for k in ref:
            d = _maxdiff(ref[k], ora[k])
            worst = max(worst, d)
            print(f"[item:{name}] {k:22s} max|ref-oracle| = {d:.3e}  |ref|max = {float(ref[k].abs().max()):.3f}")
            assert d <= TOL, (name, k, d)
        assert torch.isfinite(ref["query_outputs"]).all()
        if not args.check:
            np.savez(os.path.join(GOLDEN, f"item_{name}.npz"),
                     **{k: v.numpy() for k, v in ref.items()},
                     query_outputs_nomask_row0=ref_nomask["query_outputs"].numpy())
        del m, sd

    for name, c in USER_CASES.items():
        sd = synth.user_qformer_state_dict(**c["model"], seed=c["seed"], attn_std=c["attn_std"])
        m = User(hidden_size=c["model"]["hidden"], num_hidden_layers=c["model"]["layers"],
                 num_attention_heads=c["heads"], intermediate_size=c["model"]["inter"],
                 num_query_tokens=c["model"]["num_query"], input_embedding_dim=c["model"]["input_dim"],
                 num_item_tokens_to_predict=c["model"]["num_predict"]).eval()
        m.load_state_dict(sd, strict=True)
        x, mask = synth.user_sequences(**c["input"])
        with torch.no_grad():
            ref = m(x, mask)
            ora = O.user_qformer_forward(sd, x, mask, num_heads=c["heads"],
                                         num_item_tokens_to_predict=c["model"]["num_predict"])
        d = _maxdiff(ref, ora)
        worst = max(worst, d)
        print(f"[user:{name}] predicted_item_tokens   max|ref-oracle| = {d:.3e}  |ref|max = {float(ref.abs().max()):.3f}")
        assert d <= TOL, (name, d)
        if not args.check:
            np.savez(os.path.join(GOLDEN, f"user_{name}.npz"), predicted_item_tokens=ref.numpy())
        del m, sd

    # Positional encoding buffer (models/user_sequence_encoder.py:20-25)
    pe_ref = PE(d_model=1024).pe[:, 0, :]
    pe_ora = O.positional_encoding_table(5000, 1024)
    d = _maxdiff(pe_ref, pe_ora)
    print(f"[pe] max|ref-oracle| = {d:.3e}")
    assert d == 0.0
    if not args.check:
        np.savez(os.path.join(GOLDEN, "positional_encoding.npz"),
                 rows=pe_ref[[0, 1, 31, 32, 777, 1599]].numpy())

    # Scoring idiom: the literal per-user loop of train_item_individual_token_joint.py:405-415
    # (the file itself cannot be imported: it needs peft and calls torch.cuda.set_device(0) at
    # import, :33), candidate 0 playing "positive", the rest "negatives".
    import torch.nn.functional as F
    c = SCORING_CASE
    u = synth.normal("score_users", (c["users"], c["dim"]), c["seed"])
    C = synth.normal("score_cands", (c["cands"], c["dim"]), c["seed"])
    user_embeddings = F.normalize(u, p=2, dim=-1)
    order = []
    sims = []
    for i in range(len(user_embeddings)):
        all_items = torch.cat([F.normalize(C[:1], p=2, dim=-1), F.normalize(C[1:], p=2, dim=-1)], dim=0)
        similarities = torch.matmul(user_embeddings[i], all_items.t())
        sorted_indices = torch.argsort(similarities, descending=True)
        order.append(sorted_indices[:c["k"]])
        sims.append(similarities[sorted_indices[:c["k"]]])
    order = torch.stack(order)
    sims = torch.stack(sims)
    s_ora, i_ora = O.cosine_topk(u, C, c["k"])
    d = _maxdiff(sims, s_ora)
    print(f"[score] max|ref-oracle| score = {d:.3e}; index mismatches = {int((order != i_ora).sum())}")
    assert d <= 1e-6 and bool((order == i_ora).all())
    if not args.check:
        np.savez(os.path.join(GOLDEN, "scoring.npz"), scores=sims.numpy(), indices=order.numpy())

    print(f"oracle pinned: worst max|ref-oracle| = {worst:.3e} (tol {TOL})")


if __name__ == "__main__":
    main()
