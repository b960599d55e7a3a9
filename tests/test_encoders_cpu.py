"""CPU: oracle/encoder_oracle.py and the drop-in QFormerLoss class against tests/golden/encoders_loss.npz - outputs of
the UNMODIFIED reference classes (oracle/pin_encoders_against_reference.py: models/mwne.py TimestampEncoder /
GeoCoordinateEncoder / ImprovedMathematicalEncoder / NormalizedMathematicalEncoder, training/item_qformer_training.py
QFormerLoss)."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "encoders_loss.npz")


def _z():
    return np.load(GOLDEN)


def _sd(z, prefix):
    return {k[len(prefix) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix + ".")}


def test_event_encoder_oracle_matches_reference_golden():
    from oracle import encoder_oracle as EO
    z = _z()
    ts, coords = torch.from_numpy(z["timestamps"]), torch.from_numpy(z["coords"])
    assert float((EO.timestamp_encoder(_sd(z, "time"), ts) - torch.from_numpy(z["time_out"])).abs().max()) <= 2e-5
    assert float((EO.geo_encoder(_sd(z, "geo"), coords) - torch.from_numpy(z["geo_out"])).abs().max()) <= 2e-5
    ctx = EO.event_context(_sd(z, "time"), _sd(z, "geo"), ts.view(6, 8), coords.view(6, 8, 2))
    assert float((ctx.view(48, -1) - torch.from_numpy(z["time_out"] + z["geo_out"])).abs().max()) <= 2e-5


def test_mwne_oracle_matches_reference_golden():
    from oracle import encoder_oracle as EO
    z = _z()
    sd = _sd(z, "mwne.base_encoder")
    x = torch.from_numpy(z["numbers"])
    assert float((EO.mwne_encode(sd, x) - torch.from_numpy(z["mwne_raw_out"])).abs().max()) <= 2e-5
    got = EO.mwne_encode(sd, x, running_std=torch.from_numpy(z["mwne.normalizer.running_std"]))
    assert float((got - torch.from_numpy(z["mwne_norm_out"])).abs().max()) <= 2e-5


LOSS_KW = {"default": {}, "script": {"contrastive_weight": 0.1},
           "custom": {"reconstruction_weight": 0.7, "contrastive_weight": 0.25, "margin": 1.5}}


@pytest.mark.parametrize("tag", list(LOSS_KW))
def test_qformer_loss_class_is_a_drop_in_for_the_reference(tag):
    """Same constructor keywords / defaults, dict inputs and 3-tuple return as training/item_qformer_training.py:41-56;
    values equal to the unmodified class's on the golden inputs (and the oracle restatement agrees)."""
    from oracle import encoder_oracle as EO
    from unirec_b200.training import QFormerLoss, qformer_loss
    z = _z()
    rec, tgt, rep, pos, neg, mask = (torch.from_numpy(z[k]) for k in
                                     ("loss_rec", "loss_tgt", "loss_rep", "loss_pos", "loss_neg", "loss_mask"))
    out = {"reconstructed_fields": rec, "item_representation": rep}
    ref = z[f"loss_{tag}"]
    got = QFormerLoss(**LOSS_KW[tag])(out, {"field_embeddings": tgt}, pos, neg, mask)
    assert len(got) == 3
    for a, b in zip(got, ref):
        assert abs(float(a) - float(b)) <= 2e-6 * max(1.0, abs(float(b)))
    for a, b in zip(EO.qformer_loss(out, tgt, pos, neg, mask, **LOSS_KW[tag]), ref):
        assert abs(float(a) - float(b)) <= 2e-6 * max(1.0, abs(float(b)))
    kw = LOSS_KW[tag]
    total = qformer_loss(out, tgt, mask, pos, neg, recon_weight=kw.get("reconstruction_weight", 1.0),
                         contrastive_weight=kw.get("contrastive_weight", 0.5), margin=kw.get("margin", 0.5))
    assert abs(float(total) - float(ref[0])) <= 2e-6 * max(1.0, abs(float(ref[0])))


def test_qformer_loss_defaults_are_the_reference_defaults():
    import inspect
    from unirec_b200.training import QFormerLoss, qformer_loss
    sig = inspect.signature(QFormerLoss.__init__)
    assert [(n, p.default) for n, p in sig.parameters.items() if n != "self"] == [
        ("reconstruction_weight", 1.0), ("contrastive_weight", 0.5), ("margin", 0.5)]
    assert list(inspect.signature(QFormerLoss.forward).parameters)[1:] == [
        "model_output", "input_embeddings", "pos_rep", "neg_rep", "attention_mask"]
    assert inspect.signature(qformer_loss).parameters["contrastive_weight"].default == 0.5


def test_encoder_modules_have_the_reference_state_dict_keys():
    from unirec_b200.encoders import (GeoCoordinateEncoder, ImprovedMathematicalEncoder, NormalizedMathematicalEncoder,
                                      TimestampEncoder)
    z = _z()
    assert set(TimestampEncoder(64).state_dict()) == set(_sd(z, "time"))
    assert set(GeoCoordinateEncoder(64).state_dict()) == set(_sd(z, "geo"))
    m = NormalizedMathematicalEncoder(ImprovedMathematicalEncoder(96, 20, 50.0, True, device="cpu"))
    assert set(m.state_dict()) == set(_sd(z, "mwne"))
    m.load_state_dict(_sd(z, "mwne"), strict=True)
    with pytest.raises(ValueError):
        ImprovedMathematicalEncoder(embedding_dim=8, num_frequencies=16)
