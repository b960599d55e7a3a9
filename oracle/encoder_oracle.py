"""CPU restatement (torch fp32) of the reference's feature encoders and of QFormerLoss - TEST INFRASTRUCTURE ONLY
(only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import oracle/).

Parity status: PINNED.  oracle/pin_encoders_against_reference.py imports the unmodified models/mwne.py classes and cuts
`QFormerLoss` out of training/item_qformer_training.py, runs them on seeded inputs / weights, checks these functions
against them and stores the reference's outputs in tests/golden/encoders_loss.npz.

Each function is driven by a state dict with the reference's keys and cites the lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def timestamp_features(timestamps: torch.Tensor) -> torch.Tensor:
    """models/mwne.py:526-565: [n] -> [n, 9] (secular, then sin/cos of day, week, year, month phases)."""
    x = timestamps.float().view(-1, 1)                                   # :527
    seconds_in_year = 365.25 * 24 * 60 * 60                              # :534
    seconds_in_day = 24 * 60 * 60                                        # :539
    comps = [x / seconds_in_year]                                        # :535
    day_phase = (x % seconds_in_day) / seconds_in_day                    # :543
    comps += [torch.sin(2 * math.pi * day_phase), torch.cos(2 * math.pi * day_phase)]
    week_phase = ((x / seconds_in_day) + 4) / 7                          # :548
    comps += [torch.sin(2 * math.pi * week_phase), torch.cos(2 * math.pi * week_phase)]
    year_phase = (x % seconds_in_year) / seconds_in_year                 # :553
    comps += [torch.sin(2 * math.pi * year_phase), torch.cos(2 * math.pi * year_phase)]
    month_phase = year_phase * 12                                        # :559
    comps += [torch.sin(2 * math.pi * month_phase), torch.cos(2 * math.pi * month_phase)]
    return torch.cat(comps, dim=-1)                                      # :563


def geo_features(coordinates: torch.Tensor) -> torch.Tensor:
    """models/mwne.py:593-608: [n, 2] (lat, lon) degrees -> [n, 3] unit-sphere cartesian."""
    if coordinates.dim() != 2 or coordinates.shape[1] != 2:
        raise ValueError("Input coordinates must be of shape [batch_size, 2]")
    lat, lon = torch.deg2rad(coordinates[:, 0]), torch.deg2rad(coordinates[:, 1])
    return torch.stack([torch.cos(lat) * torch.cos(lon), torch.cos(lat) * torch.sin(lon), torch.sin(lat)], dim=-1)


def _projection(sd: Dict[str, torch.Tensor], feats: torch.Tensor) -> torch.Tensor:
    """`projection` = Linear -> GELU (exact erf) -> Linear (models/mwne.py:519-523, :584-588)."""
    h = F.gelu(F.linear(feats, sd["projection.0.weight"], sd["projection.0.bias"]))
    return F.linear(h, sd["projection.2.weight"], sd["projection.2.bias"])


def timestamp_encoder(sd, timestamps):
    return _projection(sd, timestamp_features(timestamps))               # models/mwne.py:564


def geo_encoder(sd, coordinates):
    return _projection(sd, geo_features(coordinates.float()))            # models/mwne.py:610


def event_context(sd_time, sd_geo, timestamps, coordinates):
    """models/user_sequence_encoder.py:122-127: context = time_embs + geo_embs per event; [...,] -> [..., D]."""
    lead = tuple(timestamps.shape)
    ctx = timestamp_encoder(sd_time, timestamps.reshape(-1)) + geo_encoder(sd_geo, coordinates.reshape(-1, 2))
    return ctx.view(*lead, -1)


def mwne_encode(sd: Dict[str, torch.Tensor], numbers: torch.Tensor, include_raw: bool = True,
                running_std: torch.Tensor = None, target_std: float = 1.0) -> torch.Tensor:
    """ImprovedMathematicalEncoder.forward (models/mwne.py:134-183); with running_std also the eval-mode scaling of
    MathematicallyAwareNormalizer (:55-62)."""
    shape = tuple(numbers.shape)
    x = numbers.view(-1, 1).float()
    phases = x * sd["frequencies"].unsqueeze(0)                                          # :167
    four = torch.stack([torch.cos(phases), torch.sin(phases)], dim=-1).view(x.size(0), -1)   # :169-174
    comps = [four * sd["fourier_weight"].unsqueeze(0)]                                   # :177
    if include_raw:
        comps.append(torch.cat([x, torch.sign(x)], dim=-1) * sd["raw_scale"].unsqueeze(0))   # :181-185
    if "extra_proj.weight" in sd:
        comps.append(F.linear(x, sd["extra_proj.weight"]))                               # :155-157
    emb = torch.cat(comps, dim=-1)
    if running_std is not None:
        emb = emb * torch.clamp(target_std / (running_std.unsqueeze(0) + 1e-8), min=0.1, max=10.0)   # :56-60
    return emb.view(*shape, -1)


def qformer_loss(model_output, field_embeddings, pos_rep, neg_rep, attention_mask, reconstruction_weight=1.0,
                 contrastive_weight=0.5, margin=0.5):
    """QFormerLoss.forward (training/item_qformer_training.py:41-56) -> (total, masked_recon_loss, cont_loss)."""
    unreduced = (model_output["reconstructed_fields"] - field_embeddings) ** 2                # MSELoss('none'), :49
    recon = (unreduced * attention_mask.unsqueeze(-1)).sum() / attention_mask.sum()           # :51
    # nn.TripletMarginLoss(margin): p = 2, eps = 1e-6, mean over the batch (:45, :53)
    d_ap = F.pairwise_distance(model_output["item_representation"], pos_rep, p=2.0, eps=1e-6)
    d_an = F.pairwise_distance(model_output["item_representation"], neg_rep, p=2.0, eps=1e-6)
    cont = torch.clamp(d_ap - d_an + margin, min=0.0).mean()
    return reconstruction_weight * recon + contrastive_weight * cont, recon, cont             # :54
