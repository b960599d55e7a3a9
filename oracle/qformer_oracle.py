"""CPU oracle for the nested Q-Former encode-and-rank path.  TEST INFRASTRUCTURE ONLY.

A functional fp32 restatement (plain torch ops on CPU, driven directly by a state dict)
of the reference algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product path (unirec_b200/) never
does and has no CPU fallback.

Parity status: PINNED.  oracle/pin_against_reference.py runs the UNMODIFIED reference
(/root/reference, imported through oracle/reference_shim.py) on the same synthetic weights
and inputs, checks this restatement against it (max |diff| <= 2e-5 fp32) and writes the
reference's outputs to tests/golden/*.npz.  tests/test_oracle_golden.py re-checks the oracle
against those vectors on every run.  The reference itself holds no tests or golden vectors
(SURVEY.md section 4), so the pin is "outputs of the reference itself run here".

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

LN_EPS_BERT = 1e-12  # BertConfig.layer_norm_eps default, used at models/qformer.py:65,282,368


def _linear(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _layer_norm(sd, prefix, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _split_heads(x, num_heads):
    # models/qformer.py:161-167 transpose_for_scores
    b, s, h = x.shape
    return x.view(b, s, num_heads, h // num_heads).permute(0, 2, 1, 3)


def _drop_rows(x, drop, site):
    """nn.Dropout on a [B, T, H] hidden state with the CUDA path's counter-based mask (oracle/dropout_masks.py)."""
    if drop is None:
        return x
    from . import dropout_masks as DM
    thr, seed = drop
    b, t, h = x.shape
    keep = torch.from_numpy(DM.keep_mask_rows(seed, site, b * t, h, thr)).view(b, t, h)
    return x * keep.to(x.dtype) * DM.keep_scale(thr)


def _drop_probs(probs, drop, site):
    if drop is None:
        return probs
    from . import dropout_masks as DM
    thr, seed = drop
    b, nh, nq, nk = probs.shape
    keep = torch.from_numpy(DM.keep_mask_attention(seed, site, b, nh, nq, nk, thr))
    return probs * keep.to(probs.dtype) * DM.keep_scale(thr)


def attention_block(sd: Dict[str, torch.Tensor], prefix: str, hidden, kv_source, additive_mask,
                    num_heads: int, drop=None, site_probs: int = 0, site_out: int = 0):
    """BertAttention = BertSelfAttention + BertSelfOutput (models/qformer.py:169-275, 278-289,
    322-346).  `kv_source` is `hidden` for self-attention and encoder_hidden_states for
    cross-attention (:185-198).  Dropout is identity (eval) unless `drop = (thr16, seed)` is given
    (train mode: :258 on the probabilities, :287 on the dense output)."""
    q = _split_heads(_linear(sd, prefix + ".self.query", hidden), num_heads)
    k = _split_heads(_linear(sd, prefix + ".self.key", kv_source), num_heads)
    v = _split_heads(_linear(sd, prefix + ".self.value", kv_source), num_heads)
    scores = torch.matmul(q, k.transpose(-1, -2))                      # :205
    scores = scores / math.sqrt(q.shape[-1])                           # :244 (scale BEFORE mask)
    if additive_mask is not None:
        scores = scores + additive_mask                                # :247
    probs = torch.softmax(scores, dim=-1)                              # :250
    probs = _drop_probs(probs, drop, site_probs)                       # :258
    ctx = torch.matmul(probs, v)                                       # :264
    ctx = ctx.permute(0, 2, 1, 3).contiguous()                         # :266
    ctx = ctx.view(ctx.shape[0], ctx.shape[1], -1)                     # :267-268
    out = _drop_rows(_linear(sd, prefix + ".output.dense", ctx), drop, site_out)   # :286-287
    return _layer_norm(sd, prefix + ".output.LayerNorm", out + hidden, LN_EPS_BERT)  # :288


def qformer_backbone(sd: Dict[str, torch.Tensor], prefix: str, query_embeds, encoder_hidden_states,
                     encoder_attention_mask, num_layers: int, num_heads: int, cross_freq: int, drop=None):
    """BertModel.forward in the only mode the path uses: input_ids=None, query_embeds given,
    is_decoder=False, all-ones query attention mask (models/qformer.py:804-972).

    Masks: self-attention (1-m)*-10000 with m == 1 everywhere -> zeros (:785,801);
    cross-attention PreTrainedModel.invert_attention_mask -> (1-m)*finfo(fp32).min (:927-933).
    `drop = (thr16, seed)`: train-mode dropout with the CUDA path's mask definition (dropout_masks.py).
    """
    from .dropout_masks import (KIND_CROSS_OUT, KIND_CROSS_PROBS, KIND_FFN_OUT, KIND_SELF_OUT, KIND_SELF_PROBS,
                                SITE_EMBEDDINGS, site_id)
    dtype = query_embeds.dtype
    h = _layer_norm(sd, prefix + "embeddings.LayerNorm", query_embeds, LN_EPS_BERT)  # :104-106
    h = _drop_rows(h, drop, SITE_EMBEDDINGS)                                         # :107
    b, qn, _ = h.shape
    self_mask = torch.zeros(b, 1, 1, qn, dtype=dtype)
    if encoder_attention_mask is None:
        encoder_attention_mask = torch.ones(encoder_hidden_states.shape[:2])          # :913-917
    m = encoder_attention_mask[:, None, None, :].to(dtype)
    cross_mask = (1.0 - m) * torch.finfo(dtype).min
    for i in range(num_layers):                                                        # :517
        p = f"{prefix}encoder.layer.{i}."
        h = attention_block(sd, p + "attention", h, h, self_mask, num_heads, drop,
                            site_id(i, KIND_SELF_PROBS), site_id(i, KIND_SELF_OUT))    # :417-424
        if i % cross_freq == 0:                                                        # :386-394,432
            h = attention_block(sd, p + "crossattention", h, encoder_hidden_states, cross_mask,
                                num_heads, drop, site_id(i, KIND_CROSS_PROBS),
                                site_id(i, KIND_CROSS_OUT))                            # :436-444
        inter = F.gelu(_linear(sd, p + "intermediate_query.dense", h))                 # :359-361 (erf GELU)
        out = _drop_rows(_linear(sd, p + "output_query.dense", inter), drop,
                         site_id(i, KIND_FFN_OUT))                                     # :372-373
        h = _layer_norm(sd, p + "output_query.LayerNorm", out + h, LN_EPS_BERT)        # :374, 481-484
    return h


def item_qformer_forward(sd: Dict[str, torch.Tensor], field_embeddings: torch.Tensor,
                         attention_mask: Optional[torch.Tensor] = None, num_heads: int = 16,
                         cross_freq: int = 2, drop=None) -> Dict[str, torch.Tensor]:
    """QFormerForItemRepresentation.forward (models/qformer_utils.py:37-60)."""
    b = field_embeddings.shape[0]
    num_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("qformer.encoder.layer."))
    q = sd["query_embeddings"].expand(b, -1, -1)                                       # :39
    if attention_mask is None:
        attention_mask = torch.ones(b, field_embeddings.shape[1])                     # :40-41
    out = qformer_backbone(sd, "qformer.", q, field_embeddings.float(), attention_mask,
                           num_layers, num_heads, cross_freq, drop)                    # :45-49
    rep = _linear(sd, "item_representation_head", out.mean(dim=1))                     # :50
    rec = _linear(sd, "reconstruction_head", out)                                      # :53
    rec_fields = _linear(sd, "field_projection", rec.transpose(1, 2)).transpose(1, 2)  # :54
    return {"query_outputs": out, "item_representation": rep, "reconstructed_fields": rec_fields}


def user_qformer_forward(sd: Dict[str, torch.Tensor], user_sequence_tokens: torch.Tensor,
                         attention_mask: torch.Tensor, num_heads: int = 16,
                         num_item_tokens_to_predict: int = 32) -> torch.Tensor:
    """UserQFormer.forward (training/user_qformer_training.py:47-68); cross-attention in every
    layer (:29); prediction head Linear -> GELU(erf) -> LayerNorm(eps 1e-5) -> Linear (:38-43)."""
    b = user_sequence_tokens.shape[0]
    num_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("qformer.encoder.layer."))
    q = sd["query_embeddings"].expand(b, -1, -1)
    out = qformer_backbone(sd, "qformer.", q, user_sequence_tokens.float(), attention_mask,
                           num_layers, num_heads, 1)
    rep = out.mean(dim=1)                                                              # :60
    h = F.gelu(_linear(sd, "prediction_head.0", rep))
    h = F.layer_norm(h, (h.shape[-1],), sd["prediction_head.2.weight"], sd["prediction_head.2.bias"], 1e-5)
    flat = _linear(sd, "prediction_head.3", h)                                         # :63
    return flat.view(b, num_item_tokens_to_predict, -1)                                # :64-66


def positional_encoding_table(max_len: int, d_model: int) -> torch.Tensor:
    """PositionalEncoding buffer (models/user_sequence_encoder.py:20-25), shape [max_len, d]."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def build_user_sequences(item_tokens: torch.Tensor, history: torch.Tensor, lengths: torch.Tensor,
                         context: Optional[torch.Tensor] = None):
    """Tensor part of UserSequenceEncoder.encode_user_sequence + collate padding
    (models/user_sequence_encoder.py:128-140; training/user_qformer_training.py:153-161).

    item_tokens [N, Q, D]: item query-token table (what _get_item_query_tokens_batch returns per
    item, :71-99); history [B, Hmax] item ids; lengths [B] number of valid history items;
    context [B, Hmax, D] optional time+geo embedding broadcast over the Q tokens (:130-131).
    Returns (padded [B, Hmax*Q, D], mask [B, Hmax*Q]); PE position = flattened token index
    (:136-140); padded positions are zero (:156-159).  The stray train-mode Dropout(0.1) of the
    reference's never-.eval()'d PositionalEncoding (:18,33) is identity here by design.
    """
    b, hmax = history.shape
    n, q, d = item_tokens.shape
    tok = item_tokens[history.reshape(-1)].view(b, hmax, q, d).float()
    if context is not None:
        tok = tok + context[:, :, None, :]
    seq = tok.reshape(b, hmax * q, d) + positional_encoding_table(hmax * q, d)[None]
    mask = (torch.arange(hmax * q)[None, :] < (lengths[:, None] * q)).float()
    return seq * mask[..., None], mask


def pooled_scoring_vector(tokens: torch.Tensor) -> torch.Tensor:
    """Scoring vector = mean over the token axis (the reference's own pooling,
    models/qformer_utils.py:50 / training/user_qformer_training.py:60; decision recorded in
    SURVEY.md section 8d)."""
    return tokens.float().mean(dim=1)


def cosine_topk(user_vectors: torch.Tensor, candidates: torch.Tensor, k: int):
    """Batched restatement of MRREvaluator._compute_batch_mrr's scoring
    (training/train_item_individual_token_joint.py:405-406, 412-415): L2-normalise both sides
    (F.normalize, eps 1e-12), dot product, descending order; top-k instead of a full argsort.
    Returns (scores [B,k] fp32 descending, indices [B,k] int64)."""
    u = F.normalize(user_vectors.float(), p=2, dim=-1)
    c = F.normalize(candidates.float(), p=2, dim=-1)
    sims = u @ c.t()
    return torch.topk(sims, k, dim=-1, largest=True, sorted=True)


def cosine_scores(user_vectors: torch.Tensor, candidates: torch.Tensor) -> torch.Tensor:
    u = F.normalize(user_vectors.float(), p=2, dim=-1)
    c = F.normalize(candidates.float(), p=2, dim=-1)
    return u @ c.t()
