"""Pin oracle/joint_oracle.py against the UNMODIFIED reference classes and (re)generate tests/golden/joint_scoring.npz.

Run in the authoring container only (needs /root/reference):

    python -m oracle.pin_joint_against_reference            # check + write
    python -m oracle.pin_joint_against_reference --check    # check only

training/train_item_individual_token_joint.py cannot be imported (peft is absent and the module calls
torch.cuda.set_device(0) at import time, :33), so the source text of its `InfoNCELoss` and `MRREvaluator` classes is cut
out with `ast` and executed as is in a namespace that provides the names those classes use (torch, nn, F, np, List,
tqdm, device='cpu').  No reference code is copied into this repository; only the classes' OUTPUTS are stored.
"""
from __future__ import annotations

import argparse
import ast
import os
import sys
from typing import List

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import joint_oracle as JO  # noqa: E402

REF_FILE = "/root/reference/training/train_item_individual_token_joint.py"
GOLDEN = os.path.join(ROOT, "tests", "golden", "joint_scoring.npz")
TOL = 2e-5


def load_reference_classes():
    src = open(REF_FILE).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": torch.nn, "F": torch.nn.functional, "np": np, "List": List,
          "tqdm": (lambda it, **k: it), "device": torch.device("cpu"), "ValidationDataset": object}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("InfoNCELoss", "MRREvaluator"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF_FILE, "exec"), ns)
    return ns["InfoNCELoss"], ns["MRREvaluator"]


def make_inputs():
    g = torch.Generator().manual_seed(2026)
    B, C, D = 12, 37, 256
    users = torch.randn(B, D, generator=g)
    pos = users * 0.35 + torch.randn(B, D, generator=g)            # correlated with the user: ranks vary
    negs = torch.randn(B, C, D, generator=g)
    negs[:, ::3] += users.unsqueeze(1) * torch.rand(B, (C + 2) // 3, 1, generator=g)
    lens = torch.randint(1, C + 1, (B,), generator=g)
    lens[0], lens[1] = C, 1
    masks = torch.arange(C).unsqueeze(0) < lens.unsqueeze(1)
    return users, pos, negs, masks, lens


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    InfoNCELoss, MRREvaluator = load_reference_classes()
    users, pos, negs, masks, lens = make_inputs()
    out = {"users": users.numpy(), "pos": pos.numpy(), "negs": negs.numpy(), "masks": masks.numpy(), "lens": lens.numpy()}
    worst = 0.0
    for T in (0.07, 1.0):
        crit = InfoNCELoss(temperature=T)
        ref_masked = float(crit(users, pos, negs, masks))
        ref_full = float(crit(users, pos, negs, None))
        worst = max(worst, abs(ref_masked - float(JO.infonce_loss(users, pos, negs, masks, T))),
                    abs(ref_full - float(JO.infonce_loss(users, pos, negs, None, T))))
        # per-user losses of the reference: batches of one user through the unmodified class
        per_user = torch.stack([crit(users[i:i + 1], pos[i:i + 1], negs[i:i + 1], masks[i:i + 1]) for i in range(len(users))])
        worst = max(worst, float((per_user - JO.infonce_per_user(users, pos, negs, masks, T)).abs().max()))
        out[f"loss_masked_T{T}"] = np.float32(ref_masked)
        out[f"loss_full_T{T}"] = np.float32(ref_full)
        out[f"loss_per_user_T{T}"] = per_user.numpy()

    # MRR: the reference's evaluator with a stub model that returns the user embeddings (the model call is the LLM)
    ev = MRREvaluator.__new__(MRREvaluator)
    ev.model = lambda **kw: users
    neg_list = [negs[i, :int(lens[i])] for i in range(len(users))]
    batch = {"input_ids": torch.zeros(len(users), 1, dtype=torch.long), "attention_mask": torch.ones(len(users), 1),
             "history_field_embeddings": torch.zeros(len(users), 1, 1, 1), "history_attention_mask": torch.ones(len(users), 1, 1),
             "positive_item_embeddings": pos, "negative_item_embeddings": neg_list}
    with torch.no_grad():
        ref_mrr = ev._compute_batch_mrr(batch)
    assert ref_mrr == JO.reciprocal_ranks(users, pos, neg_list), (ref_mrr, JO.reciprocal_ranks(users, pos, neg_list))
    out["mrr"] = np.asarray(ref_mrr, dtype=np.float64)
    print(f"oracle vs reference classes: max |diff| = {worst:.2e}; reciprocal ranks identical: {ref_mrr}")
    assert worst <= TOL, worst
    if not args.check:
        np.savez_compressed(GOLDEN, **out)
        print("wrote", GOLDEN)


if __name__ == "__main__":
    main()
