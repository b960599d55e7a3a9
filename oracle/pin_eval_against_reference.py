"""Pin oracle/eval_oracle.py against the UNMODIFIED reference `evaluate_reconstruction_quality`
(evaluation/evaluate_item_qformer.py:40-103) and (re)generate tests/golden/eval_metrics.npz.

Run in the authoring container only (needs /root/reference):   python -m oracle.pin_eval_against_reference [--check]

The reference function is imported through oracle/reference_shim.py and run on CPU with the reference model holding
the synthetic weights of the 'small' golden case, over a temporary field-embedding cache in the reference's own format
(embeddings.pt / masks.pt dictionaries).  Only its two result numbers are stored.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import eval_oracle as EO  # noqa: E402
from oracle import qformer_oracle as O  # noqa: E402
from oracle import reference_shim  # noqa: E402
from tests.golden_cases import EVAL_CASE, ITEM_CASES  # noqa: E402
from unirec_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "eval_metrics.npz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    Item, _, _ = reference_shim.load()
    from evaluation.evaluate_item_qformer import evaluate_reconstruction_quality
    c = ITEM_CASES[EVAL_CASE["item_case"]]
    mk = c["model"]
    sd = synth.item_qformer_state_dict(**mk, seed=c["seed"], attn_std=c["attn_std"])
    m = Item(hidden_size=mk["hidden"], num_hidden_layers=mk["layers"], num_attention_heads=c["heads"],
             intermediate_size=mk["inter"], num_query_tokens=mk["num_query"], field_embedding_dim=mk["field_dim"],
             num_fields=mk["num_fields"]).eval()
    m.load_state_dict(sd, strict=True)
    x, mask = synth.item_fields(**EVAL_CASE["input"])
    with tempfile.TemporaryDirectory() as d:
        torch.save({i: x[i] for i in range(len(x))}, os.path.join(d, "embeddings.pt"))
        torch.save({i: mask[i] for i in range(len(x))}, os.path.join(d, "masks.pt"))
        with contextlib.redirect_stdout(io.StringIO()):
            ref = evaluate_reconstruction_quality(m, d, device="cpu", batch_size=EVAL_CASE["batch_size"])
    with torch.no_grad():
        got = EO.reconstruction_quality(
            lambda f, mm: O.item_qformer_forward(sd, f, mm, num_heads=c["heads"])["reconstructed_fields"], x, mask,
            EVAL_CASE["batch_size"])
    print("reference:", ref, "\noracle   :", got)
    for k in ref:
        assert abs(ref[k] - got[k]) <= 2e-5 * max(1.0, abs(ref[k])), (k, ref[k], got[k])
    if not args.check:
        np.savez_compressed(GOLDEN, val_recon_loss=np.float64(ref["val_recon_loss"]),
                            avg_cosine_similarity=np.float64(ref["avg_cosine_similarity"]))
        print("wrote", GOLDEN)


if __name__ == "__main__":
    main()
