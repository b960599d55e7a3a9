"""CPU: the oracle (oracle/qformer_oracle.py) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/*.npz, generator: oracle/pin_against_reference.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import qformer_oracle as O
from tests.golden_cases import ITEM_CASES, SCORING_CASE, USER_CASES
from unirec_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5  # fp32 CPU vs fp32 CPU, same torch build


@pytest.mark.parametrize("name", list(ITEM_CASES))
def test_item_oracle_matches_reference_golden(name):
    c = ITEM_CASES[name]
    sd = synth.item_qformer_state_dict(**c["model"], seed=c["seed"], attn_std=c["attn_std"])
    x, mask = synth.item_fields(**c["input"])
    out = O.item_qformer_forward(sd, x, mask, num_heads=c["heads"])
    g = np.load(os.path.join(GOLDEN, f"item_{name}.npz"))
    for k in ("query_outputs", "item_representation", "reconstructed_fields"):
        assert np.abs(out[k].numpy() - g[k]).max() <= TOL, k
    nomask = O.item_qformer_forward(sd, x[:1], None, num_heads=c["heads"])["query_outputs"]
    assert np.abs(nomask.numpy() - g["query_outputs_nomask_row0"]).max() <= TOL
    assert np.isfinite(g["query_outputs"]).all()  # includes the all-masked row


@pytest.mark.parametrize("name", list(USER_CASES))
def test_user_oracle_matches_reference_golden(name):
    c = USER_CASES[name]
    sd = synth.user_qformer_state_dict(**c["model"], seed=c["seed"], attn_std=c["attn_std"])
    x, mask = synth.user_sequences(**c["input"])
    out = O.user_qformer_forward(sd, x, mask, num_heads=c["heads"],
                                 num_item_tokens_to_predict=c["model"]["num_predict"])
    g = np.load(os.path.join(GOLDEN, f"user_{name}.npz"))
    assert np.abs(out.numpy() - g["predicted_item_tokens"]).max() <= TOL


def test_positional_encoding_golden():
    g = np.load(os.path.join(GOLDEN, "positional_encoding.npz"))["rows"]
    pe = O.positional_encoding_table(1600, 1024)[[0, 1, 31, 32, 777, 1599]].numpy()
    assert np.abs(pe - g).max() == 0.0


def test_scoring_golden():
    c = SCORING_CASE
    u = synth.normal("score_users", (c["users"], c["dim"]), c["seed"])
    C = synth.normal("score_cands", (c["cands"], c["dim"]), c["seed"])
    s, i = O.cosine_topk(u, C, c["k"])
    g = np.load(os.path.join(GOLDEN, "scoring.npz"))
    assert np.abs(s.numpy() - g["scores"]).max() <= 1e-6
    assert (i.numpy() == g["indices"]).all()


def test_all_masked_item_is_uniform_attention():
    """an item whose fields are all masked gives finite output equal to
    what uniform attention over the (zero) field vectors gives, i.e. context = value bias."""
    c = ITEM_CASES["small"]
    sd = synth.item_qformer_state_dict(**c["model"], seed=c["seed"], attn_std=c["attn_std"])
    x = torch.zeros(1, 6, 256)
    out_masked = O.item_qformer_forward(sd, x, torch.zeros(1, 6), num_heads=4)["query_outputs"]
    out_open = O.item_qformer_forward(sd, x, torch.ones(1, 6), num_heads=4)["query_outputs"]
    assert torch.isfinite(out_masked).all()
    assert (out_masked - out_open).abs().max() <= 1e-5


def test_bf16_precision_model_brackets_the_tolerances():
    """The CPU precision model (bf16 rounding at the kernels' storage points) meets the standard bar on
    the reference-like and 2-layer sharp cases, and shows that the 12-layer sharp case is chaotic in
    bf16 (which is why tests/test_modules_gpu.py bounds that case relative to this model)."""
    from oracle import bf16_precision_model as P
    stats = {}
    for name in ("sharp2", "small", "sharp"):
        c = ITEM_CASES[name]
        sd = synth.item_qformer_state_dict(**c["model"], seed=c["seed"], attn_std=c["attn_std"])
        x, mask = synth.item_fields(**c["input"])
        g = np.load(os.path.join(GOLDEN, f"item_{name}.npz"))["query_outputs"]
        stats[name] = P.error_stats(P.item_query_outputs(sd, x, mask, num_heads=c["heads"]), torch.as_tensor(g))
    for name in ("sharp2", "small"):
        mx, mean, cos = stats[name]
        assert mx <= 0.15 and mean <= 0.02 and cos >= 0.9995, (name, stats[name])
    assert stats["sharp"][0] > 0.15  # documents the amplification; not a bound on the kernels
