"""CPU precision model of the CUDA path.  TEST INFRASTRUCTURE ONLY (same import rule as
oracle/qformer_oracle.py: tests/, smoke() and bench.py's cpu_baseline only).

Same algorithm as oracle/qformer_oracle.py (which follows models/qformer.py:103-108, 169-289,
349-375, 402-484 of the reference), evaluated in fp32 on CPU but ROUNDED TO bf16 AT THE POINTS
WHERE THE KERNELS STORE bf16: weights, projection outputs, softmax probabilities (the P operand
of the PV mma), attention context, GELU output and every LayerNorm output except the last.
Accumulation, softmax and LayerNorm statistics stay fp32, exactly as in the kernels.

Purpose: it tells the tests how far ANY bf16-storage implementation of this network is from
the fp32 reference on a given input.  The network is a 12-layer post-LN transformer; with
sharp (near one-hot) softmax, a bf16 rounding of a logit can flip an attention winner and the
difference is amplified layer after layer, so the tolerance of SURVEY.md section 8c is stated as
"no worse than 1.5x the bf16 error measured on the same inputs" - this file measures that error.
It is not a second oracle: parity is always judged against the fp32 reference outputs.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

LN_EPS_BERT = 1e-12


def _bf(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).float()


def _lin(sd, p, x):
    return F.linear(x, _bf(sd[p + ".weight"]), sd[p + ".bias"])


def _ln(sd, p, x, eps=LN_EPS_BERT):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _heads(x, nh):
    b, s, h = x.shape
    return x.view(b, s, nh, h // nh).permute(0, 2, 1, 3)


def _attention(sd, p, h, kv, add_mask, nh, pre_ln_bf16):
    q = _heads(_bf(_lin(sd, p + ".self.query", h)), nh)
    k = _heads(_bf(_lin(sd, p + ".self.key", kv)), nh)
    v = _heads(_bf(_lin(sd, p + ".self.value", kv)), nh)
    s = q @ k.transpose(-1, -2) / 8.0
    if add_mask is not None:
        s = s + add_mask
    pr = torch.softmax(s, dim=-1)
    ctx = _bf(pr) @ v
    ctx = _bf(ctx.permute(0, 2, 1, 3).reshape(h.shape[0], h.shape[1], -1))
    pre = _lin(sd, p + ".output.dense", ctx) + h
    if pre_ln_bf16:
        pre = _bf(pre)
    return _bf(_ln(sd, p + ".output.LayerNorm", pre))


def backbone(sd: Dict[str, torch.Tensor], query_embeddings: torch.Tensor, enc: torch.Tensor,
             mask: Optional[torch.Tensor], num_heads: int, cross_freq: int, pre_ln_bf16: bool = False):
    b = enc.shape[0]
    layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("qformer.encoder.layer."))
    h = _bf(_ln(sd, "qformer.embeddings.LayerNorm", query_embeddings.expand(b, -1, -1)))
    enc = _bf(enc.float())
    cm = None
    if mask is not None:
        cm = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    for i in range(layers):
        p = f"qformer.encoder.layer.{i}."
        h = _attention(sd, p + "attention", h, h, None, num_heads, pre_ln_bf16)
        if i % cross_freq == 0:
            h = _attention(sd, p + "crossattention", h, enc, cm, num_heads, pre_ln_bf16)
        inter = _bf(F.gelu(_lin(sd, p + "intermediate_query.dense", h)))
        pre = _lin(sd, p + "output_query.dense", inter) + h
        if pre_ln_bf16:
            pre = _bf(pre)
        h = _ln(sd, p + "output_query.LayerNorm", pre)
        if i < layers - 1:
            h = _bf(h)
    return h


@torch.no_grad()
def item_query_outputs(sd, field_embeddings, attention_mask, num_heads=16, pre_ln_bf16=False):
    return backbone(sd, sd["query_embeddings"], field_embeddings, attention_mask, num_heads, 2, pre_ln_bf16)


@torch.no_grad()
def user_last_hidden(sd, user_sequence_tokens, attention_mask, num_heads=16, pre_ln_bf16=False):
    return backbone(sd, sd["query_embeddings"], user_sequence_tokens, attention_mask, num_heads, 1, pre_ln_bf16)


def error_stats(out: torch.Tensor, ref: torch.Tensor):
    """(max|d|, mean|d|, cosine) of two tensors, the three figures the parity tests bound."""
    o, r = out.float().flatten(), ref.float().flatten()
    d = (o - r).abs()
    return float(d.max()), float(d.mean()), float(F.cosine_similarity(o, r, dim=0))
