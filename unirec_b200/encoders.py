"""Drop-in mirrors of the reference's event / number feature encoders on the sm_100a kernels.

Same class names, constructor arguments, parameter / buffer names (state-dict keys) and forward contracts as
  * TimestampEncoder             - models/mwne.py:504-566
  * GeoCoordinateEncoder         - models/mwne.py:569-610
  * ImprovedMathematicalEncoder  - models/mwne.py:91-183
  * NormalizedMathematicalEncoder (eval mode) over MathematicallyAwareNormalizer - models/mwne.py:9-62, :186-222
plus `encode_event_context`, the fused form of what UserSequenceEncoder.encode_user_sequence does with the first two
(models/user_sequence_encoder.py:117-131: context = time_emb + geo_emb, added to every query token of the event) - one
feature kernel and ONE tcgen05 GEMM for both encoders, producing the `context` tensor `ops.build_user_sequence` /
`NestedRanker` consume.  No CPU fallback: inputs must be CUDA tensors.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import ops


def _mlp(in_dim: int, embedding_dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_dim, embedding_dim * 2), nn.GELU(), nn.Linear(embedding_dim * 2, embedding_dim))


class _EventEncoder(nn.Module):
    """Shared body: `projection` = Linear(in, 2D) -> GELU -> Linear(2D, D) exactly as the reference declares it."""
    IN_DIM = 0

    def __init__(self, embedding_dim: int):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.projection = _mlp(self.IN_DIM, embedding_dim)
        self._pack = None
        self._pack_key = None

    def packed(self) -> dict:
        ps = [self.projection[0].weight, self.projection[0].bias, self.projection[2].weight, self.projection[2].bias]
        key = (ps[0].device, tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))
        if self._pack is None or self._pack_key != key:
            self._pack = {"w1": ps[0].detach().float().contiguous(), "b1": ps[1].detach().float().contiguous(),
                          "w2": ps[2].detach().to(torch.bfloat16).contiguous(), "b2": ps[3].detach().float().contiguous()}
            self._pack_key = key
        return self._pack

    def invalidate_packed(self):
        self._pack = self._pack_key = None

    def _second_layer(self, hidden_half: torch.Tensor) -> torch.Tensor:
        pk = self.packed()
        return ops.linear(hidden_half, pk["w2"], pk["b2"], out_dtype=torch.float32)


class TimestampEncoder(_EventEncoder):
    """models/mwne.py:504-566.  forward(timestamps [n] Unix seconds, any real dtype) -> fp32 [n, embedding_dim]."""
    IN_DIM = 9

    @torch.no_grad()
    def forward(self, timestamps: torch.Tensor) -> torch.Tensor:
        pk = self.packed()
        H2 = 2 * self.embedding_dim
        hid = ops.context_hidden(timestamps.reshape(-1), None, pk["w1"], pk["b1"], None, None, H2)
        return self._second_layer(hid[:, :H2])


class GeoCoordinateEncoder(_EventEncoder):
    """models/mwne.py:569-610.  forward(coordinates [n, 2] (lat, lon) degrees) -> fp32 [n, embedding_dim]."""
    IN_DIM = 3

    @torch.no_grad()
    def forward(self, coordinates: torch.Tensor) -> torch.Tensor:
        if coordinates.dim() != 2 or coordinates.shape[1] != 2:
            raise ValueError("Input coordinates must be of shape [batch_size, 2]")
        pk = self.packed()
        H2 = 2 * self.embedding_dim
        hid = ops.context_hidden(None, coordinates.float(), None, None, pk["w1"], pk["b1"], H2)
        return self._second_layer(hid[:, H2:])


@torch.no_grad()
def encode_event_context(timestamp_encoder: TimestampEncoder, geo_encoder: GeoCoordinateEncoder,
                         timestamps: torch.Tensor, coordinates: torch.Tensor,
                         out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """context = timestamp_encoder(timestamps) + geo_encoder(coordinates) (models/user_sequence_encoder.py:122-127) for
    every event of every user: timestamps [B, Hmax] (or [n]), coordinates [B, Hmax, 2] (or [n, 2]) ->
    [B, Hmax, D] (or [n, D]) in `out_dtype`.  One feature kernel + one GEMM with K = 4 D:
        [gelu(W1t f_t + b1t) | gelu(W1g f_g + b1g)] x [W2t | W2g]^T + (b2t + b2g)."""
    D = timestamp_encoder.embedding_dim
    if geo_encoder.embedding_dim != D:
        raise ValueError("timestamp and geo encoders must share embedding_dim")
    lead = tuple(timestamps.shape)
    if tuple(coordinates.shape) != lead + (2,):
        raise ValueError(f"coordinates must be {lead + (2,)}, got {tuple(coordinates.shape)}")
    pt, pg = timestamp_encoder.packed(), geo_encoder.packed()
    key = (timestamp_encoder._pack_key, geo_encoder._pack_key)
    cat = getattr(timestamp_encoder, "_cat", None)
    if cat is None or cat[0] != key:
        cat = (key, torch.cat([pt["w2"], pg["w2"]], 1).contiguous(), (pt["b2"] + pg["b2"]).contiguous())
        timestamp_encoder._cat = cat
    hid = ops.context_hidden(timestamps.reshape(-1), coordinates.reshape(-1, 2).float(), pt["w1"], pt["b1"], pg["w1"],
                             pg["b1"], 2 * D)
    return ops.linear(hid, cat[1], cat[2], out_dtype=out_dtype).view(*lead, D)


class ImprovedMathematicalEncoder(nn.Module):
    """models/mwne.py:91-183: Fourier features at log-spaced frequencies + raw value / sign + a learned linear part."""

    def __init__(self, embedding_dim: int = 64, num_frequencies: int = 16, max_frequency: float = 100.0,
                 include_raw: bool = True, device: str = "cuda"):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.num_frequencies = num_frequencies
        self.max_frequency = max_frequency
        self.include_raw = include_raw
        self.device = device
        self.register_buffer("frequencies", torch.logspace(-2, math.log10(max_frequency), num_frequencies))
        fourier_dim = 2 * num_frequencies
        raw_dim = 2 if include_raw else 0
        remaining_dim = embedding_dim - fourier_dim - raw_dim
        if remaining_dim < 0:
            raise ValueError(f"embedding_dim {embedding_dim} too small for {fourier_dim} + {raw_dim}")
        self.fourier_weight = nn.Parameter(torch.ones(fourier_dim))
        self.extra_proj = nn.Linear(1, remaining_dim, bias=False) if remaining_dim > 0 else None
        if include_raw:
            self.raw_scale = nn.Parameter(torch.tensor([1.0, 1.0]))

    @torch.no_grad()
    def forward(self, numbers: torch.Tensor, _scale: Optional[torch.Tensor] = None) -> torch.Tensor:
        shape = tuple(numbers.shape)
        x = numbers.reshape(-1).float()
        out = ops.mwne_encode(x, self.frequencies.float().contiguous(), self.fourier_weight.detach().float().contiguous(),
                              self.raw_scale.detach().float().contiguous() if self.include_raw else None,
                              self.extra_proj.weight.detach().float().reshape(-1).contiguous()
                              if self.extra_proj is not None else None, _scale, self.embedding_dim)
        return out.view(*shape, self.embedding_dim)


class MathematicallyAwareNormalizer(nn.Module):
    """Buffers of models/mwne.py:9-28 (so reference checkpoints load); eval-mode forward only (:55-62)."""

    def __init__(self, encoder, target_std=1.0, momentum=0.99, min_std=0.1):
        super().__init__()
        self.encoder = encoder
        self.target_std, self.momentum, self.min_std = target_std, momentum, min_std
        self.register_buffer("running_mean", torch.zeros(encoder.embedding_dim))
        self.register_buffer("running_std", torch.ones(encoder.embedding_dim))
        self.register_buffer("num_batches_tracked", torch.tensor(0))
        self.register_buffer("additivity_error_history", torch.zeros(100))
        self.register_buffer("history_idx", torch.tensor(0))

    @torch.no_grad()
    def forward(self, numbers):
        if self.training:
            raise NotImplementedError("unirec_b200: the normaliser's running-statistics update (train mode) is not on the "
                                      "inference path; call .eval() (the reference loads it with .eval(), mwne.py:643)")
        scale = torch.clamp(self.target_std / (self.running_std.float() + 1e-8), min=0.1, max=10.0).contiguous()
        return self.encoder(numbers, _scale=scale)


class NormalizedMathematicalEncoder(nn.Module):
    """models/mwne.py:186-222 (what `load_trained_encoder` returns, :612-660, and ItemEncoder uses for numeric fields)."""

    def __init__(self, base_encoder: ImprovedMathematicalEncoder, target_std: float = 1.0, momentum: float = 0.99,
                 min_std: float = 0.1):
        super().__init__()
        self.base_encoder = base_encoder
        self.normalizer = MathematicallyAwareNormalizer(base_encoder, target_std, momentum, min_std)

    def forward(self, numbers: torch.Tensor) -> torch.Tensor:
        return self.normalizer(numbers)

    @property
    def embedding_dim(self):
        return self.base_encoder.embedding_dim

    @property
    def num_frequencies(self):
        return self.base_encoder.num_frequencies

    @property
    def max_frequency(self):
        return self.base_encoder.max_frequency

    @property
    def include_raw(self):
        return self.base_encoder.include_raw
