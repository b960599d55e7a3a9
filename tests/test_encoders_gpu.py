"""GPU parity of the feature-encoder kernels (csrc/encoders.cu through the C ABI) against the golden outputs of the
UNMODIFIED reference classes (tests/golden/encoders_loss.npz) and against the CPU oracle on seeded inputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "encoders_loss.npz")


def _sd(z, prefix):
    return {k[len(prefix) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix + ".")}


def _close(got, ref, cos_min=0.9995, rel_max=0.02):
    got, ref = got.float().cpu(), ref.float()
    cos = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0))
    rel = float((got - ref).abs().max() / ref.abs().max())
    assert cos >= cos_min and rel <= rel_max, (cos, rel)


def test_event_features_match_reference_to_a_few_ulp():
    """The 9 + 3 raw features in fp32: same operation order as models/mwne.py:526-565, :596-608; sin / cos of arguments
    up to ~2e4 rad agree with torch's CPU sinf within 2e-6 absolute."""
    from unirec_b200 import ops
    z = np.load(GOLDEN)
    ts, coords = torch.from_numpy(z["timestamps"]), torch.from_numpy(z["coords"])
    sd_t, sd_g = _sd(z, "time"), _sd(z, "geo")
    for tsv in (ts.to(DEV), ts.float().to(DEV), ts.double().to(DEV)):           # int64, fp32 and fp64 inputs
        _, feats = ops.context_hidden(tsv, coords.to(DEV), sd_t["projection.0.weight"].to(DEV),
                                      sd_t["projection.0.bias"].to(DEV), sd_g["projection.0.weight"].to(DEV),
                                      sd_g["projection.0.bias"].to(DEV), 128, want_features=True)
        f = feats.cpu()
        assert float((f[:, :9] - torch.from_numpy(z["time_feats"])).abs().max()) <= 2e-6
        assert float((f[:, 9:] - torch.from_numpy(z["geo_feats"])).abs().max()) <= 2e-6


def test_timestamp_and_geo_encoders_match_reference_golden():
    """Tolerance: the second Linear runs on the bf16 tensor cores (fp32 accumulate) - cosine >= 0.9995 and
    max|d| <= 2 % of the output range against the reference's fp32 outputs."""
    from unirec_b200.encoders import GeoCoordinateEncoder, TimestampEncoder, encode_event_context
    z = np.load(GOLDEN)
    te, ge = TimestampEncoder(64), GeoCoordinateEncoder(64)
    te.load_state_dict(_sd(z, "time"), strict=True)
    ge.load_state_dict(_sd(z, "geo"), strict=True)
    te, ge = te.to(DEV).eval(), ge.to(DEV).eval()
    ts, coords = torch.from_numpy(z["timestamps"]).to(DEV), torch.from_numpy(z["coords"]).to(DEV)
    out_t, out_g = te(ts), ge(coords)
    assert out_t.dtype == torch.float32 and tuple(out_t.shape) == (48, 64)
    _close(out_t, torch.from_numpy(z["time_out"]))
    _close(out_g, torch.from_numpy(z["geo_out"]))
    ctx = encode_event_context(te, ge, ts.view(6, 8), coords.view(6, 8, 2), out_dtype=torch.float32)
    assert tuple(ctx.shape) == (6, 8, 64)
    _close(ctx.view(48, 64), torch.from_numpy(z["time_out"] + z["geo_out"]))
    with pytest.raises(ValueError):
        ge(coords.view(-1))                              # models/mwne.py:593-594
    with pytest.raises(RuntimeError):
        te(ts.cpu())                                     # no CPU path


def test_mwne_encoder_matches_reference_golden():
    from unirec_b200.encoders import ImprovedMathematicalEncoder, NormalizedMathematicalEncoder
    z = np.load(GOLDEN)
    m = NormalizedMathematicalEncoder(ImprovedMathematicalEncoder(96, 20, 50.0, True, device=DEV))
    m.load_state_dict(_sd(z, "mwne"), strict=True)
    m = m.to(DEV).eval()
    x = torch.from_numpy(z["numbers"]).to(DEV)
    raw, norm = m.base_encoder(x).cpu(), m(x).cpu()
    ref_raw, ref_norm = torch.from_numpy(z["mwne_raw_out"]), torch.from_numpy(z["mwne_norm_out"])
    # fp32 elementwise: only sinf / cosf of |x f| up to 5e5 rad differ from torch's CPU implementation, by ulps of 1
    assert float((raw - ref_raw).abs().max()) <= 1e-5 * float(ref_raw.abs().max())
    assert float((norm - ref_norm).abs().max()) <= 1e-5 * float(ref_norm.abs().max())
    assert tuple(m(x.view(2, 13)).shape) == (2, 13, 96)


def test_event_context_full_width_feeds_the_sequence_builder():
    """D = 1024 (the path's width): context for [B, Hmax] events -> `ops.build_user_sequence(..., context)` equals the
    oracle's tokens + context + PE (models/user_sequence_encoder.py:122-140)."""
    from oracle import encoder_oracle as EO
    from oracle import qformer_oracle as O
    from unirec_b200 import ops
    from unirec_b200.encoders import GeoCoordinateEncoder, TimestampEncoder, encode_event_context
    torch.manual_seed(3)
    D, B, Hmax, Q, N = 1024, 5, 7, 32, 40
    te, ge = TimestampEncoder(D).to(DEV).eval(), GeoCoordinateEncoder(D).to(DEV).eval()
    g = torch.Generator().manual_seed(4)
    ts = torch.randint(1_500_000_000, 1_760_000_000, (B, Hmax), generator=g)
    coords = torch.stack([torch.rand(B, Hmax, generator=g) * 180 - 90, torch.rand(B, Hmax, generator=g) * 360 - 180], -1)
    ctx = encode_event_context(te, ge, ts.to(DEV), coords.to(DEV))
    ref_ctx = EO.event_context({k: v.float().cpu() for k, v in te.state_dict().items()},
                               {k: v.float().cpu() for k, v in ge.state_dict().items()}, ts, coords)
    _close(ctx, ref_ctx)
    table = (torch.randn(N, Q, D, generator=g) * 0.5).to(torch.bfloat16)
    hist = torch.randint(0, N, (B, Hmax), generator=g)
    lens = torch.tensor([7, 1, 3, 7, 5], dtype=torch.int32)
    seq, mask = ops.build_user_sequence(table.to(DEV), hist.to(DEV), lens.to(DEV), ctx)
    ref_seq, ref_mask = O.build_user_sequences(table.float(), hist, lens.long(), context=ctx.float().cpu())
    assert torch.equal(mask.cpu(), ref_mask.float())
    # the context reaches the builder as bf16 and the sum is stored as bf16: two roundings of 2^-9 relative each
    err = (seq.float().cpu() - ref_seq).abs()
    assert bool((err <= 2 * 2.0 ** -8 * ref_seq.abs() + 1e-3).all()), float(err.max())
