"""Hooks of the reference's joint trainer that sit on either side of the Q-Former path, on the
sm_100a kernels.  The LLM (Qwen3 + LoRA), its tokenizer and the HF Trainer stay what they are - out of scope; what the
joint trainer does AROUND them with Q-Former tensors is here, under the reference's names:

  * `history_query_tokens`   - item Q-Former over the [B, num_hist, F, D] history fields
                               (training/train_item_individual_token_joint.py:146-157)
  * `inject_history_tokens`  - the placeholder-token overwrite of the text embeddings (:160-171: a triple Python loop
                               with one `nonzero` per (item, token, batch element)) as ONE kernel
  * `InfoNCELoss`            - same constructor / forward signature as :326-352 (padded negatives + masks), one
                               streaming kernel + one per-user reduction instead of normalize x3, bmm, masked_fill
                               and a Python loop over the batch; differentiable (autograd.Function, backward kernel)
  * `batch_mrr`              - the scoring half of MRREvaluator._compute_batch_mrr (:405-418) for ragged
                               per-user negative lists, without the per-user Python loop

No CPU fallback: inputs must be CUDA tensors.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import torch
import torch.nn as nn

from . import ops


def history_query_tokens(qformer_model, history_field_embeddings: torch.Tensor,
                         history_attention_mask: torch.Tensor) -> torch.Tensor:
    """[B, num_hist, F, D] fields + [B, num_hist, F] mask -> [B, num_hist, Q, H] query tokens (:146-157)."""
    bh, num_hist, num_fields, field_dim = history_field_embeddings.shape
    out = qformer_model(history_field_embeddings.reshape(bh * num_hist, num_fields, field_dim),
                        history_attention_mask.reshape(bh * num_hist, num_fields))["query_outputs"]
    return out.view(bh, num_hist, out.shape[1], out.shape[2])


class _InjectFn(torch.autograd.Function):
    """The reference's overwrite `text_embeds[b, positions] = query_embeddings[b]` (:160-171) is a differentiable
    index assignment: it is the only path by which the joint trainer's loss reaches the item Q-Former.  Forward = the
    in-place kernel (text_embeds is marked dirty, so autograd sees the version bump); backward = one kernel that moves the
    gradient of every overwritten position to the token that replaced it and zeroes it for the text embedding."""

    @staticmethod
    def forward(ctx, text_embeds, input_ids, token_ids, tokens):
        ctx.mark_dirty(text_embeds)
        ctx.save_for_backward(input_ids, token_ids)
        ctx.tok_dtype = tokens.dtype
        ops.inject_tokens(text_embeds, input_ids, token_ids, tokens.detach())
        return text_embeds

    @staticmethod
    def backward(ctx, d_out):
        input_ids, token_ids = ctx.saved_tensors
        d_text = d_out.contiguous().clone()
        d_tokens = ops.inject_tokens_backward(d_text, input_ids, token_ids)
        return (d_text if ctx.needs_input_grad[0] else None), None, None, \
            (d_tokens.to(ctx.tok_dtype) if ctx.needs_input_grad[3] else None)


def inject_history_tokens(text_embeds: torch.Tensor, input_ids: torch.Tensor, token_ids: torch.Tensor,
                          history_item_query_tokens: torch.Tensor) -> torch.Tensor:
    """text_embeds[b, positions of <|history_item_i_query_j|>] = history_item_query_tokens[b, i, j] for every (i, j)
    (:160-171), in place and differentiable: gradients reach `history_item_query_tokens` (hence the item Q-Former) and
    stop at the overwritten rows of `text_embeds`, exactly like the reference's indexed assignment.
    token_ids int64 [num_hist, Q] (or flat): tokenizer ids of the placeholder tokens."""
    B, nh, Q, Hd = history_item_query_tokens.shape
    toks = history_item_query_tokens.reshape(B, nh * Q, Hd)
    tids = token_ids.reshape(-1).to(text_embeds.device)
    if torch.is_grad_enabled() and (text_embeds.requires_grad or toks.requires_grad):
        return _InjectFn.apply(text_embeds, input_ids, tids, toks)
    return ops.inject_tokens(text_embeds, input_ids, tids, toks)


class _InfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, users, pos, negs, mask, temperature):
        u = users.detach().contiguous()
        p = pos.detach().to(u.dtype).contiguous()
        n = negs.detach().to(u.dtype).contiguous()
        sims, inv = ops.list_scores(u, p, n, mask=mask)
        loss, _ = ops.infonce_rank(sims, temperature)
        ctx.save_for_backward(u, p, n, sims, inv)
        ctx.mask, ctx.temperature = mask, temperature
        ctx.dtypes = (users.dtype, pos.dtype, negs.dtype)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        u, p, n, sims, inv = ctx.saved_tensors
        want_list = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        d_user, d_list = ops.list_scores_backward(u, p, n, sims, inv, dloss.float().contiguous(), ctx.temperature,
                                                  mask=ctx.mask, want_list_grad=want_list)
        du = d_user.to(ctx.dtypes[0]) if ctx.needs_input_grad[0] else None
        dp = d_list[:, 0].to(ctx.dtypes[1]) if ctx.needs_input_grad[1] else None
        dn = d_list[:, 1:].to(ctx.dtypes[2]) if ctx.needs_input_grad[2] else None
        return du, dp, dn, None, None


class InfoNCELoss(nn.Module):
    """Drop-in for the reference's InfoNCELoss (training/train_item_individual_token_joint.py:326-352)."""

    def __init__(self, temperature: float = 0.07):
        super().__init__()
        self.temperature = temperature

    def per_user(self, user_embeddings, positive_item_embeddings, negative_item_embeddings, negative_masks=None):
        if not user_embeddings.is_cuda:
            raise RuntimeError("unirec_b200.InfoNCELoss: inputs must be CUDA tensors (no CPU fallback)")
        mask = None if negative_masks is None else negative_masks.to(torch.bool)
        return _InfoNCEFn.apply(user_embeddings, positive_item_embeddings, negative_item_embeddings, mask,
                                float(self.temperature))

    def forward(self, user_embeddings, positive_item_embeddings, negative_item_embeddings, negative_masks=None):
        return self.per_user(user_embeddings, positive_item_embeddings, negative_item_embeddings, negative_masks).mean()


@torch.no_grad()
def positive_ranks(user_embeddings: torch.Tensor, positive_item_embeddings: torch.Tensor,
                   negative_item_embeddings: Union[torch.Tensor, Sequence[torch.Tensor]]) -> torch.Tensor:
    """1-based rank of every user's positive among [positive] + its negatives by cosine similarity (int32 [B]).
    negative_item_embeddings: a list of [n_b, D] tensors (the validation collate, :381-390) or a padded [B, C, D]."""
    u = user_embeddings.contiguous()
    p = positive_item_embeddings.to(u.dtype).contiguous()
    if isinstance(negative_item_embeddings, torch.Tensor):
        sims, _ = ops.list_scores(u, p, negative_item_embeddings.to(u.dtype).contiguous())
    else:
        lens = [int(t.shape[0]) for t in negative_item_embeddings]
        offs = torch.zeros(len(lens) + 1, dtype=torch.int64)
        offs[1:] = torch.tensor(lens, dtype=torch.int64).cumsum(0)
        flat = torch.cat([t.to(device=u.device, dtype=u.dtype) for t in negative_item_embeddings], 0).contiguous()
        sims, _ = ops.list_scores(u, p, flat, offsets=offs.to(u.device), max_list=max(lens) if lens else 0)
    return ops.infonce_rank(sims, 1.0)[1]


def batch_mrr(user_embeddings, positive_item_embeddings, negative_item_embeddings) -> List[float]:
    """Reciprocal ranks of one validation batch, as MRREvaluator._compute_batch_mrr returns them (:405-418)."""
    return (1.0 / positive_ranks(user_embeddings, positive_item_embeddings, negative_item_embeddings).float()).tolist()
