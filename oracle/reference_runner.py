"""Drives the UNMODIFIED reference modules (imported through oracle/reference_shim.py from /root/reference or its verbatim
copy oracle/_ref/) over the path's workload - TEST / BASELINE INFRASTRUCTURE ONLY (bench.py `--impl reference`,
bench.py's same-GPU eager comparator, tests).  Nothing here is the product path.

What is the reference's own code here: QFormerForItemRepresentation, UserQFormer and PositionalEncoding (constructed by
their own constructors, weights loaded with load_state_dict(strict=True)).  What is glue, following the cited lines:
  * per-user sequence assembly = the tensor half of UserSequenceEncoder.encode_user_sequence
    (models/user_sequence_encoder.py:128-140: flatten, PositionalEncoding(x.unsqueeze(1)).squeeze(1)) - its item-side
    half (feature encoders on raw JSON items) is outside the path, the item tokens come from the resident table;
  * padding = collate_fn (training/user_qformer_training.py:153-161);
  * scoring = F.normalize both sides, matmul, descending order (training/train_item_individual_token_joint.py:405-415)
    with the pooled vectors of SURVEY.md 8d.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import reference_shim


def available() -> bool:
    return reference_shim.available()


def load_models(item_sd: Optional[dict], user_sd: Optional[dict], device="cpu", num_fields: int = 14, d_model: int = 1024,
                item_kwargs: Optional[dict] = None, user_kwargs: Optional[dict] = None):
    """Reference modules in eval() with the given state dicts (strict).  Returns (item or None, user or None, PE)."""
    Item, User, PE = reference_shim.load()
    item = user = None
    with torch.no_grad():
        if item_sd is not None:
            item = Item(num_fields=num_fields, **(item_kwargs or {}))
            item.load_state_dict(item_sd, strict=True)
            item = item.to(device).eval()
        if user_sd is not None:
            user = User(**(user_kwargs or {}))
            user.load_state_dict(user_sd, strict=True)
            user = user.to(device).eval()
        pe = PE(d_model=d_model).to(device).eval()          # eval(): its stray Dropout(0.1) is identity (see the oracle)
    return item, user, pe


@torch.no_grad()
def item_tokens(item, fields: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """data_processing/qformer_inference.py:160-163: model(batch) -> query_outputs."""
    return item(fields, mask)["query_outputs"]


@torch.no_grad()
def user_sequences(pe, tokens_per_user: torch.Tensor, lengths: torch.Tensor):
    """tokens_per_user [B, Hmax, Q, D] (history items' query tokens), lengths [B] -> (padded [B, Hmax*Q, D], mask)."""
    B, Hmax, Q, D = tokens_per_user.shape
    encoded = []
    for b in range(B):
        n = int(lengths[b])
        flat = tokens_per_user[b, :n].reshape(n * Q, D)                         # user_sequence_encoder.py:133-136
        encoded.append(pe(flat.unsqueeze(1)).squeeze(1) if n > 0 else flat)     # :139-140
    max_len = Hmax * Q
    padded = torch.zeros(B, max_len, D, device=tokens_per_user.device)          # user_qformer_training.py:154-155
    mask = torch.zeros(B, max_len, device=tokens_per_user.device)
    for i, seq in enumerate(encoded):
        padded[i, :seq.shape[0]] = seq                                          # :158-160
        mask[i, :seq.shape[0]] = 1
    return padded, mask


@torch.no_grad()
def user_rank(user, pe, tokens_per_user: torch.Tensor, lengths: torch.Tensor, candidates: torch.Tensor, k: int):
    """One pass of the nested user path through the reference modules -> (scores [B, k], indices [B, k])."""
    seq, mask = user_sequences(pe, tokens_per_user, lengths)
    pred = user(seq, mask)                                                      # UserQFormer.forward
    u = F.normalize(pred.mean(dim=1), p=2, dim=-1)                              # pooled scoring vector
    c = F.normalize(candidates, p=2, dim=-1)                                    # joint trainer :405-406, :412
    sims = torch.matmul(u, c.t())                                               # :413
    return torch.topk(sims, k, dim=-1)                                          # top of the descending order (:414)
