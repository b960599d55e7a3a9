"""CPU: the joint-trainer oracle (oracle/joint_oracle.py) against the golden vectors produced by the UNMODIFIED
reference classes (oracle/pin_joint_against_reference.py -> tests/golden/joint_scoring.npz)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "joint_scoring.npz")


def _load():
    z = np.load(GOLDEN)
    return z, [torch.from_numpy(z[k]) for k in ("users", "pos", "negs", "masks", "lens")]


def test_infonce_oracle_matches_reference_golden():
    from oracle import joint_oracle as JO
    z, (users, pos, negs, masks, lens) = _load()
    for T in (0.07, 1.0):
        assert abs(float(JO.infonce_loss(users, pos, negs, masks, T)) - float(z[f"loss_masked_T{T}"])) <= 2e-5
        assert abs(float(JO.infonce_loss(users, pos, negs, None, T)) - float(z[f"loss_full_T{T}"])) <= 2e-5
        per = JO.infonce_per_user(users, pos, negs, masks, T)
        assert float((per - torch.from_numpy(z[f"loss_per_user_T{T}"])).abs().max()) <= 2e-5


def test_mrr_oracle_matches_reference_golden():
    from oracle import joint_oracle as JO
    z, (users, pos, negs, masks, lens) = _load()
    neg_list = [negs[i, :int(lens[i])] for i in range(len(users))]
    assert JO.reciprocal_ranks(users, pos, neg_list) == z["mrr"].tolist()
    assert len(set(z["mrr"].tolist())) > 3          # the fixture exercises several different ranks


def test_inject_tokens_oracle_semantics():
    from oracle import joint_oracle as JO
    B, S, nh, Q, Hd = 2, 9, 2, 3, 8
    ids = torch.arange(100, 100 + nh * Q).view(nh, Q)
    input_ids = torch.tensor([[1, 100, 101, 2, 105, 105, 3, 4, 5], [104, 1, 1, 1, 1, 1, 1, 1, 102]])
    text = torch.zeros(B, S, Hd)
    toks = torch.arange(B * nh * Q * Hd, dtype=torch.float32).view(B, nh, Q, Hd)
    out = JO.inject_tokens(text, input_ids, ids, toks)
    assert torch.equal(out[0, 1], toks[0, 0, 0]) and torch.equal(out[0, 2], toks[0, 0, 1])
    assert torch.equal(out[0, 4], toks[0, 1, 2]) and torch.equal(out[0, 5], toks[0, 1, 2])     # every occurrence
    assert torch.equal(out[1, 0], toks[1, 1, 1]) and torch.equal(out[1, 8], toks[1, 0, 2])
    assert float(out[0, 0].abs().sum()) == 0 and float(out[1, 1:8].abs().sum()) == 0


def test_inject_tokens_oracle_matches_reference_forward_golden():
    """tests/golden/joint_scoring.npz `inj_*`: the embeddings the (stub) LLM received from the UNMODIFIED
    MultiModalQwenEmbedding.forward (training/train_item_individual_token_joint.py:134-181) - the oracle's loop must
    reproduce them bit for bit."""
    from oracle import joint_oracle as JO
    z = np.load(GOLDEN)
    got = JO.inject_tokens(torch.from_numpy(z["inj_text"]), torch.from_numpy(z["inj_input_ids"]),
                           torch.from_numpy(z["inj_token_ids"]), torch.from_numpy(z["inj_tokens"]))
    assert torch.equal(got, torch.from_numpy(z["inj_out"]))
    assert not torch.equal(got, torch.from_numpy(z["inj_text"]))


def test_eval_oracle_matches_reference_golden():
    """oracle/eval_oracle.py over the CPU item-Q-Former oracle reproduces the two numbers the UNMODIFIED reference
    function evaluate_reconstruction_quality returned (oracle/pin_eval_against_reference.py)."""
    from oracle import eval_oracle as EO
    from oracle import qformer_oracle as O
    from tests.golden_cases import EVAL_CASE, ITEM_CASES
    from unirec_b200 import synth
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_metrics.npz"))
    c = ITEM_CASES[EVAL_CASE["item_case"]]
    sd = synth.item_qformer_state_dict(**c["model"], seed=c["seed"], attn_std=c["attn_std"])
    x, mask = synth.item_fields(**EVAL_CASE["input"])
    with torch.no_grad():
        got = EO.reconstruction_quality(
            lambda f, m: O.item_qformer_forward(sd, f, m, num_heads=c["heads"])["reconstructed_fields"], x, mask,
            EVAL_CASE["batch_size"])
    assert abs(got["val_recon_loss"] - float(z["val_recon_loss"])) <= 2e-5 * abs(float(z["val_recon_loss"]))
    assert abs(got["avg_cosine_similarity"] - float(z["avg_cosine_similarity"])) <= 2e-6
