"""Pin oracle/joint_oracle.py against the UNMODIFIED reference classes and (re)generate tests/golden/joint_scoring.npz.

Run in the authoring container only (needs /root/reference):

    python -m oracle.pin_joint_against_reference            # check + write
    python -m oracle.pin_joint_against_reference --check    # check only

training/train_item_individual_token_joint.py cannot be imported (peft is absent and the module calls
torch.cuda.set_device(0) at import time, :33), so the source text of its `InfoNCELoss`, `MRREvaluator` and
`MultiModalQwenEmbedding` classes is cut out with `ast` and executed as is in a namespace that provides the names those classes use (torch, nn, F, np, List,
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


def load_reference_joint_model_class():
    """`MultiModalQwenEmbedding` (:88-181) cut out unmodified; its __init__ (Qwen3 download, LoRA) is never called - the
    pin builds the object with __new__ and gives it stub collaborators, then runs the UNMODIFIED forward (:134-181)."""
    src = open(REF_FILE).read()
    tree = ast.parse(src)
    from typing import Optional
    ns = {"torch": torch, "nn": torch.nn, "F": torch.nn.functional, "np": np, "List": List, "Optional": Optional,
          "LoraConfig": object, "device": torch.device("cpu"), "os": os}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "MultiModalQwenEmbedding":
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF_FILE, "exec"), ns)
    return ns["MultiModalQwenEmbedding"]


class _StubTokenizer:
    def __init__(self, ids):
        self.ids = ids

    def convert_tokens_to_ids(self, name):
        return self.ids[name]


class _StubLLM(torch.nn.Module):
    """Stands in for Qwen3: an embedding table + an identity 'transformer' that records the embeddings it was given."""

    def __init__(self, vocab, hidden, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.emb = torch.nn.Embedding(vocab, hidden)
        self.emb.weight.data = torch.randn(vocab, hidden, generator=g)
        self.seen = None

    def get_input_embeddings(self):
        return self.emb

    def forward(self, inputs_embeds=None, attention_mask=None, output_hidden_states=True):
        self.seen = inputs_embeds.detach().clone()
        from types import SimpleNamespace
        return SimpleNamespace(hidden_states=[inputs_embeds])


def injection_case():
    """Seeded inputs of the token-injection pin: every placeholder once, three of them twice, one user missing four."""
    g = torch.Generator().manual_seed(77)
    B, S, nh, Q, Hd, vocab = 4, 160, 5, 8, 32, 400
    token_ids = (300 + torch.randperm(nh * Q, generator=g)).view(nh, Q)
    input_ids = torch.randint(0, 290, (B, S), generator=g)
    for b in range(B):
        perm = torch.randperm(S, generator=g)[: nh * Q + 3]
        input_ids[b, perm[: nh * Q]] = token_ids.reshape(-1)
        input_ids[b, perm[nh * Q:]] = token_ids.reshape(-1)[:3]
        if b == 2:
            input_ids[b, perm[5:9]] = 11
    tokens = torch.randn(B, nh, Q, Hd, generator=g)
    return B, S, nh, Q, Hd, vocab, token_ids, input_ids, tokens


def pin_injection():
    """Runs the reference's forward with stubs and returns (inputs, the embeddings the LLM received)."""
    B, S, nh, Q, Hd, vocab, token_ids, input_ids, tokens = injection_case()
    Model = load_reference_joint_model_class()
    m = Model.__new__(Model)
    torch.nn.Module.__init__(m)
    m.base_model = _StubLLM(vocab, Hd, seed=5)
    m.tokenizer = _StubTokenizer({f"<|history_item_{i}_query_{j}|>": int(token_ids[i, j]) for i in range(nh) for j in range(Q)})
    m.num_history_items, m.num_query_tokens_per_item = nh, Q
    m.qformer_model = lambda x, mask: {"query_outputs": tokens.view(B * nh, Q, Hd)}
    with torch.no_grad():
        pooled = m.forward(input_ids, torch.ones(B, S), torch.zeros(B, nh, 1, 1), torch.ones(B, nh, 1))
    seen = m.base_model.seen
    text = m.base_model.emb.weight.data[input_ids]
    mine = JO.inject_tokens(text, input_ids, token_ids, tokens)
    assert torch.equal(mine, seen), "oracle inject_tokens differs from the reference forward"
    assert torch.equal(pooled, seen.mean(dim=1))
    return {"inj_token_ids": token_ids.numpy(), "inj_input_ids": input_ids.numpy(), "inj_tokens": tokens.numpy(),
            "inj_text": text.numpy(), "inj_out": seen.numpy()}


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
    out.update(pin_injection())
    print("token injection: oracle == unmodified MultiModalQwenEmbedding.forward (stub LLM / tokenizer), bit-exact")
    print(f"oracle vs reference classes: max |diff| = {worst:.2e}; reciprocal ranks identical: {ref_mrr}")
    assert worst <= TOL, worst
    if not args.check:
        np.savez_compressed(GOLDEN, **out)
        print("wrote", GOLDEN)


if __name__ == "__main__":
    main()
