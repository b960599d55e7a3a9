"""Pin oracle/encoder_oracle.py against the UNMODIFIED reference and (re)generate tests/golden/encoders_loss.npz.

Run in the authoring container only (needs /root/reference):
    python -m oracle.pin_encoders_against_reference            # check + write
    python -m oracle.pin_encoders_against_reference --check    # check only

models/mwne.py imports cleanly (torch / numpy only): TimestampEncoder, GeoCoordinateEncoder, ImprovedMathematicalEncoder
and NormalizedMathematicalEncoder are the reference's own classes.  `QFormerLoss` is cut out of
training/item_qformer_training.py with `ast` (the module pulls in the feature encoders at import).  Only inputs, seeded
weights and the reference's OUTPUTS are stored.
"""
from __future__ import annotations

import argparse
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import encoder_oracle as EO  # noqa: E402

REF_ROOT = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden", "encoders_loss.npz")
D = 64          # embedding_dim of the pinned encoders (the kernels are dimension-agnostic; 2D = 128 = two K blocks)


def cases():
    g = torch.Generator().manual_seed(515)
    n = 48
    ts = torch.randint(1_200_000_000, 1_760_000_000, (n,), generator=g)          # 2008 .. 2025, int64 like torch.tensor([...])
    ts[0], ts[1], ts[2] = 0, 86_399, 1_700_000_000
    coords = torch.stack([torch.rand(n, generator=g) * 180 - 90, torch.rand(n, generator=g) * 360 - 180], 1)
    coords[0] = torch.tensor([90.0, 0.0])
    coords[1] = torch.tensor([-90.0, 180.0])
    numbers = torch.cat([torch.randn(20, generator=g) * 50, torch.tensor([0.0, -0.0, 1.0, -3.5, 1e4, 19.99])])
    return ts, coords, numbers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    sys.path.insert(0, REF_ROOT)
    from models.mwne import (GeoCoordinateEncoder, ImprovedMathematicalEncoder, NormalizedMathematicalEncoder,
                             TimestampEncoder)
    ts, coords, numbers = cases()
    out = {"timestamps": ts.numpy(), "coords": coords.numpy(), "numbers": numbers.numpy()}
    worst = 0.0
    with torch.no_grad():
        torch.manual_seed(11)
        te = TimestampEncoder(D).eval()
        ge = GeoCoordinateEncoder(D).eval()
        for name, m in (("time", te), ("geo", ge)):
            for k, v in m.state_dict().items():
                out[f"{name}.{k}"] = v.numpy()
        ref_t, ref_g = te(ts), ge(coords)
        sd_t = {k: v for k, v in te.state_dict().items()}
        sd_g = {k: v for k, v in ge.state_dict().items()}
        worst = max(worst, float((EO.timestamp_encoder(sd_t, ts) - ref_t).abs().max()),
                    float((EO.geo_encoder(sd_g, coords) - ref_g).abs().max()))
        out["time_out"], out["geo_out"] = ref_t.numpy(), ref_g.numpy()
        out["time_feats"] = EO.timestamp_features(ts).numpy()      # oracle features (== what the reference projected)
        out["geo_feats"] = EO.geo_features(coords).numpy()
        # the features ARE the reference's: its projection applied to them must give its output
        assert torch.equal(te.projection(EO.timestamp_features(ts)), ref_t)
        assert torch.equal(ge.projection(EO.geo_features(coords)), ref_g)

        torch.manual_seed(12)
        base = ImprovedMathematicalEncoder(embedding_dim=96, num_frequencies=20, max_frequency=50.0, include_raw=True,
                                           device="cpu")
        base.fourier_weight.data = torch.rand(40) + 0.5
        base.raw_scale.data = torch.tensor([0.3, 1.7])
        norm = NormalizedMathematicalEncoder(base, target_std=1.0).eval()
        norm.normalizer.running_std = torch.rand(96) * 12 + 0.05      # exercises both clamps of the scale
        for k, v in norm.state_dict().items():
            out[f"mwne.{k}"] = v.numpy()
        ref_raw, ref_norm = base(numbers), norm(numbers)
        sd_b = {k: v for k, v in base.state_dict().items()}
        worst = max(worst, float((EO.mwne_encode(sd_b, numbers) - ref_raw).abs().max()),
                    float((EO.mwne_encode(sd_b, numbers, running_std=norm.normalizer.running_std) - ref_norm).abs().max()))
        out["mwne_raw_out"], out["mwne_norm_out"] = ref_raw.numpy(), ref_norm.numpy()

    # QFormerLoss: the unmodified class cut out of the training script
    src = open(os.path.join(REF_ROOT, "training/item_qformer_training.py")).read()
    ns = {"torch": torch, "nn": torch.nn}
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == "QFormerLoss":
            exec(compile(ast.Module(body=[node], type_ignores=[]), "item_qformer_training.py", "exec"), ns)
    QFormerLoss = ns["QFormerLoss"]
    g = torch.Generator().manual_seed(99)
    B, F_, E = 24, 6, 32
    rec = torch.randn(B, F_, E, generator=g)
    tgt = torch.randn(B, F_, E, generator=g)
    rep = torch.randn(B, E, generator=g)
    pos = rep * 0.6 + torch.randn(B, E, generator=g) * 0.5
    neg = torch.randn(B, E, generator=g)
    mask = (torch.rand(B, F_, generator=g) < 0.75).long()
    mask[0] = 1
    out.update({"loss_rec": rec.numpy(), "loss_tgt": tgt.numpy(), "loss_rep": rep.numpy(), "loss_pos": pos.numpy(),
                "loss_neg": neg.numpy(), "loss_mask": mask.numpy()})
    for tag, kw in (("default", {}), ("script", {"contrastive_weight": 0.1}),
                    ("custom", {"reconstruction_weight": 0.7, "contrastive_weight": 0.25, "margin": 1.5})):
        crit = QFormerLoss(**kw)
        ref = crit({"reconstructed_fields": rec, "item_representation": rep}, {"field_embeddings": tgt}, pos, neg, mask)
        mine = EO.qformer_loss({"reconstructed_fields": rec, "item_representation": rep}, tgt, pos, neg, mask, **kw)
        worst = max(worst, max(abs(float(a) - float(b)) for a, b in zip(ref, mine)))
        out[f"loss_{tag}"] = np.asarray([float(v) for v in ref], dtype=np.float64)
    print(f"oracle vs reference (encoders, MWNE, QFormerLoss): max |diff| = {worst:.2e}")
    assert worst <= 2e-5, worst
    if not args.check:
        np.savez_compressed(GOLDEN, **out)
        print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    main()
