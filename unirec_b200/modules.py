"""Drop-in nn.Module mirrors of the reference's Q-Former wrappers, running on the sm_100a kernels.

Same class names, constructor arguments, attributes (`.config`, `.num_query_tokens`, `.qformer`,
`.query_embeddings`), output contract and **state-dict keys** as
  * QFormerForItemRepresentation  - models/qformer_utils.py:16-60 (identical copy models/qformer_model.py:6-50)
  * UserQFormer                   - training/user_qformer_training.py:17-68
so that `load_state_dict(checkpoint['model_state_dict'])` of a reference checkpoint works unchanged,
including the tensors the query-only path never executes (word/position embeddings, the text-branch
FFN `intermediate`/`output`).

The forward pass is NOT the reference's eager op sequence.  Per call it runs, per layer:
  one fused QKV projection GEMM (tcgen05) -> fused self-attention -> output-projection GEMM with the
  residual added in the epilogue -> LayerNorm; cross-attention K/V for ALL cross layers come from ONE
  GEMM over the field/sequence embeddings (the encoder input is layer-invariant); the FFN is a
  GEMM+bias+erf-GELU and a GEMM+bias+residual, then LayerNorm.  The query-token LayerNorm is
  batch-invariant and is computed on Q rows and broadcast.  Activations are bf16, accumulation and all
  softmax / LayerNorm statistics are fp32.

There is no CPU fallback: inputs must be CUDA tensors and the module must live on a CUDA device.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops

LN_EPS = 1e-12  # BertConfig.layer_norm_eps default used by the reference backbone (models/qformer.py:65)


def _bert_config(**kw):
    from transformers.models.bert.configuration_bert import BertConfig
    return BertConfig(**kw)


# ------------------------------------------------------------------------------------------------
# Parameter containers with the reference's attribute names (no compute in these classes).
# ------------------------------------------------------------------------------------------------
class _SelfAttentionParams(nn.Module):
    def __init__(self, hidden: int, kv_in: int):
        super().__init__()
        self.query = nn.Linear(hidden, hidden)
        self.key = nn.Linear(kv_in, hidden)
        self.value = nn.Linear(kv_in, hidden)


class _DenseLN(nn.Module):
    def __init__(self, fan_in: int, hidden: int, eps: float):
        super().__init__()
        self.dense = nn.Linear(fan_in, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)


class _AttentionParams(nn.Module):
    def __init__(self, hidden: int, kv_in: int, eps: float):
        super().__init__()
        setattr(self, "self", _SelfAttentionParams(hidden, kv_in))
        self.output = _DenseLN(hidden, hidden, eps)


class _Dense(nn.Module):
    def __init__(self, fan_in: int, fan_out: int):
        super().__init__()
        self.dense = nn.Linear(fan_in, fan_out)


class _LayerParams(nn.Module):
    def __init__(self, cfg, layer_num: int):
        super().__init__()
        h, i = cfg.hidden_size, cfg.intermediate_size
        self.attention = _AttentionParams(h, h, cfg.layer_norm_eps)
        self.has_cross_attention = bool(cfg.add_cross_attention and layer_num % cfg.cross_attention_freq == 0)
        if self.has_cross_attention:
            self.crossattention = _AttentionParams(h, cfg.encoder_width, cfg.layer_norm_eps)
        # text branch: allocated and checkpointed by the reference, never executed on this path
        self.intermediate = _Dense(h, i)
        self.output = _DenseLN(i, h, cfg.layer_norm_eps)
        self.intermediate_query = _Dense(h, i)
        self.output_query = _DenseLN(i, h, cfg.layer_norm_eps)


class _EncoderParams(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([_LayerParams(cfg, i) for i in range(cfg.num_hidden_layers)])


class _EmbeddingParams(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size, padding_idx=cfg.pad_token_id)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(cfg.max_position_embeddings).expand((1, -1)))


class QFormerBackbone(nn.Module):
    """Parameter tree of the reference `BertModel(config, add_pooling_layer=False)`
    (models/qformer.py:677-697) plus the fused CUDA forward of its query-only mode (:804-972)."""

    def __init__(self, cfg):
        super().__init__()
        if cfg.hidden_size % cfg.num_attention_heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (cfg.hidden_size, cfg.num_attention_heads))  # models/qformer.py:115-121
        if cfg.hidden_size // cfg.num_attention_heads != 64:
            raise ValueError("unirec_b200 attention kernels are built for head_dim 64 "
                             f"(got {cfg.hidden_size // cfg.num_attention_heads})")
        self.config = cfg
        self.embeddings = _EmbeddingParams(cfg)
        self.encoder = _EncoderParams(cfg)
        self._init_weights()
        self._pack: Optional[dict] = None
        self._pack_key = None
        self.hoist_layer0 = True        # compute the batch-invariant head of the encoder once (see `encode`)
        # True: cross-attention over long key sequences (64 queries, S % 64 == 0) projects K / V inside the attention
        # kernel (csrc/kv_attention_fused.cu) instead of materialising them; UserQFormer switches it on
        self.fused_kv_attention = False
        # Long key sequences (the user model): the K/V of ONE cross-attention layer are materialised for at most this many
        # bytes at a time (a chunk of users) while every other op of the layer runs on ALL users of the call - see
        # `encode`.  None: K/V of all layers and all users in one GEMM (the item model: 14 keys per item).
        self.kv_chunk_bytes: Optional[int] = None
        # True: no LayerNorm kernel and no LayerNorm output between the GEMMs of the encoder - the dense + residual GEMM that
        # writes a LayerNorm's input also emits its row statistics, and the GEMMs that consume the LayerNorm read that input
        # with the normalisation folded into their weights / epilogues (`_encode_from_kv_folded`, ops.linear_ln)
        # (effective with bf16 pre-LayerNorm buffers and hidden / FFN widths that are multiples of 256; measured +5.3 % items/s,
        # +0.4..2 % users/s on one box, profiles/r02_l_*)
        self.fold_layernorm = True
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    def _init_weights(self):
        # reference rule, models/qformer.py:664-674: N(0, initializer_range) weights, zero bias, LN (1, 0)
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data.normal_(mean=0.0, std=std)
            elif isinstance(m, nn.LayerNorm):
                m.bias.data.zero_()
                m.weight.data.fill_(1.0)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()

    # -------------------------------------------------------------------------------- weight packing
    def _live_params(self) -> List[torch.Tensor]:
        ps = [self.embeddings.LayerNorm.weight, self.embeddings.LayerNorm.bias]
        for lyr in self.encoder.layer:
            blocks = [lyr.attention] + ([lyr.crossattention] if lyr.has_cross_attention else [])
            for blk in blocks:
                s = getattr(blk, "self")
                ps += [s.query.weight, s.query.bias, s.key.weight, s.key.bias, s.value.weight, s.value.bias,
                       blk.output.dense.weight, blk.output.dense.bias, blk.output.LayerNorm.weight,
                       blk.output.LayerNorm.bias]
            ps += [lyr.intermediate_query.dense.weight, lyr.intermediate_query.dense.bias,
                   lyr.output_query.dense.weight, lyr.output_query.dense.bias,
                   lyr.output_query.LayerNorm.weight, lyr.output_query.LayerNorm.bias]
        return ps

    def invalidate_packed(self):
        """Drop the cached bf16 weight pack.  The cache key is the parameters' autograd version counters, which in-place
        writes through `.data` do NOT bump (`dist.broadcast(p.data, 0)`, `p.data.copy_()`, EMA weight swaps): call this
        after such an update.  `load_state_dict` and `.to()` / `.half()`-style `_apply` calls do it automatically."""
        self._pack = self._pack_key = None

    def _apply(self, fn, *a, **k):
        self.invalidate_packed()
        return super()._apply(fn, *a, **k)

    def packed(self) -> dict:
        """bf16 / fused copies of the live weights, rebuilt only when a parameter changed (see `invalidate_packed`)."""
        live = self._live_params()
        key = (live[0].device, tuple(p._version for p in live), tuple(p.data_ptr() for p in live))
        if self._pack is not None and self._pack_key == key:
            return self._pack
        bf = torch.bfloat16
        f32 = lambda t: t.detach().float().contiguous()
        layers = []
        kv_w, kv_b = [], []
        h_ = self.config.hidden_size
        for lyr in self.encoder.layer:
            a = getattr(lyr.attention, "self")
            d = {
                "w_qkv": torch.cat([a.query.weight, a.key.weight, a.value.weight], 0).detach().to(bf).contiguous(),
                "b_qkv": torch.cat([a.query.bias, a.key.bias, a.value.bias], 0).detach().float().contiguous(),
                "w_o": lyr.attention.output.dense.weight.detach().to(bf).contiguous(),
                "b_o": f32(lyr.attention.output.dense.bias),
                "ln1_g": f32(lyr.attention.output.LayerNorm.weight), "ln1_b": f32(lyr.attention.output.LayerNorm.bias),
                "w_1": lyr.intermediate_query.dense.weight.detach().to(bf).contiguous(),
                "b_1": f32(lyr.intermediate_query.dense.bias),
                "w_2": lyr.output_query.dense.weight.detach().to(bf).contiguous(),
                "b_2": f32(lyr.output_query.dense.bias),
                "ln3_g": f32(lyr.output_query.LayerNorm.weight), "ln3_b": f32(lyr.output_query.LayerNorm.bias),
                "cross": lyr.has_cross_attention,
            }
            if lyr.has_cross_attention:
                c = getattr(lyr.crossattention, "self")
                d.update({
                    "w_qc": c.query.weight.detach().to(bf).contiguous(), "b_qc": f32(c.query.bias),
                    "w_oc": lyr.crossattention.output.dense.weight.detach().to(bf).contiguous(),
                    "b_oc": f32(lyr.crossattention.output.dense.bias),
                    "ln2_g": f32(lyr.crossattention.output.LayerNorm.weight),
                    "ln2_b": f32(lyr.crossattention.output.LayerNorm.bias),
                    "kv_slot": len(kv_w),
                    "b_v": f32(c.value.bias),
                })
                if h_ % 128 == 0:        # an even number of heads: operand of the fused K/V-projection + attention kernel
                    d["w_kvp"] = ops.pack_kv_weights(c.key.weight, c.value.weight)
                kv_w.append(torch.cat([c.key.weight, c.value.weight], 0))
                kv_b.append(torch.cat([c.key.bias, c.value.bias], 0))
            layers.append(d)
        pack = {
            "layers": layers,
            "w_kv_all": torch.cat(kv_w, 0).detach().to(bf).contiguous() if kv_w else None,
            "b_kv_all": torch.cat(kv_b, 0).detach().float().contiguous() if kv_b else None,
            "emb_g": f32(self.embeddings.LayerNorm.weight), "emb_b": f32(self.embeddings.LayerNorm.bias),
        }
        self._pack, self._pack_key = pack, key
        return pack

    def _layer0_invariants(self, pk: dict, query_embeddings: torch.Tensor, prelayernorm_dtype: torch.dtype) -> dict:
        """The batch-invariant head of the encoder on Q rows: pre1 = dense(self-attention(LN(queries))) + LN(queries)
        (the input of layer 0's first LayerNorm) and qc = layer 0's cross-attention query projection of LN1(pre1).
        Cached in the weight pack (rebuilt with it) and keyed by the query tensor's version."""
        key = (query_embeddings.data_ptr(), query_embeddings._version, prelayernorm_dtype)
        inv = pk.get("l0")
        if inv is not None and inv["key"] == key:
            return inv
        cfg = self.config
        H, heads = cfg.hidden_size, cfg.num_attention_heads
        Q = query_embeddings.shape[1]
        L = pk["layers"][0]
        q0 = query_embeddings.detach().reshape(Q, H).float().contiguous()
        h0 = ops.layernorm(q0, pk["emb_g"], pk["emb_b"], cfg.layer_norm_eps)
        qkv = ops.linear(h0, L["w_qkv"], L["b_qkv"])
        ctx = ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], batch=1, num_heads=heads, nq=Q, nk=Q)
        pre1 = ops.linear(ctx, L["w_o"], L["b_o"], epilogue=ops.EPI_BIAS_RESIDUAL, residual=h0,
                          out_dtype=prelayernorm_dtype)
        h1 = ops.layernorm(pre1, L["ln1_g"], L["ln1_b"], cfg.layer_norm_eps)
        inv = {"key": key, "pre1": pre1, "qc": ops.linear(h1, L["w_qc"], L["b_qc"])}
        pk["l0"] = inv
        return inv

    # -------------------------------------------------------------------------------------- forward
    def fused_kv_supported(self, S: int) -> bool:
        """The fused K/V-projection + cross-attention kernel (csrc/kv_attention_fused.cu) covers this model / sequence."""
        cfg = self.config
        return (self.fused_kv_attention and not self.training and
                ops.kv_attention_supported(cfg.query_length, S, cfg.num_attention_heads, cfg.encoder_width))

    def encode(self, query_embeddings: torch.Tensor, encoder_hidden_states: torch.Tensor,
               encoder_attention_mask: Optional[torch.Tensor], out_dtype: torch.dtype = torch.float32,
               prelayernorm_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """query_embeddings [1, Q, H] (learned tokens, fp32); encoder_hidden_states [B, S, E] (bf16 or
        fp32); encoder_attention_mask [B, S] (1 attend / 0 masked) or None.  Returns last_hidden_state
        [B, Q, H] in `out_dtype` (models/qformer.py:957)."""
        cfg = self.config
        if not encoder_hidden_states.is_cuda:
            raise RuntimeError("unirec_b200: encoder_hidden_states must be a CUDA tensor (no CPU fallback)")
        if encoder_attention_mask is not None and encoder_attention_mask.dim() != 2:
            raise ValueError("Wrong shape for encoder_attention_mask (shape {})".format(
                tuple(encoder_attention_mask.shape)))
        B, S, E = encoder_hidden_states.shape
        H, heads = cfg.hidden_size, cfg.num_attention_heads
        Q = query_embeddings.shape[1]
        if E != cfg.encoder_width:
            raise ValueError(f"encoder width {E} != config.encoder_width {cfg.encoder_width}")
        pk = self.packed()

        enc = ops.cast_bf16(encoder_hidden_states.contiguous()).view(B * S, E)
        mask = None
        if encoder_attention_mask is not None:
            mask = encoder_attention_mask.to(device=enc.device, dtype=torch.float32).contiguous()
        if Q == cfg.query_length and self.fused_kv_supported(S):
            # long key sequences (the user model): every cross-attention layer projects its K/V tile by tile INSIDE the
            # attention kernel - no K/V buffer (13.4 GB per 512 users) is written to or read from HBM
            return self.encode_from_kv(query_embeddings, None, B, S, mask, out_dtype, prelayernorm_dtype, enc=enc)
        n_cross = sum(1 for L in pk["layers"] if L["cross"])
        if self.kv_chunk_bytes is not None and n_cross and B * S * 2 * H * n_cross * 2 > self.kv_chunk_bytes:
            # layer-major: K/V of one layer for a chunk of users at a time, everything else on all B users at once
            return self.encode_from_kv(query_embeddings, None, B, S, mask, out_dtype, prelayernorm_dtype, enc=enc,
                                       kv_chunk_users=self.kv_chunk_users(B, S))
        # cross-attention K/V of every cross layer in one GEMM (the encoder input is layer-invariant)
        kv_all = ops.linear(enc, pk["w_kv_all"], pk["b_kv_all"]) if pk["w_kv_all"] is not None else None
        return self.encode_from_kv(query_embeddings, kv_all, B, S, mask, out_dtype, prelayernorm_dtype)

    def kv_chunk_users(self, B: int, S: int) -> int:
        """Users per K/V chunk of the layer-major path: as many as fit `kv_chunk_bytes` for one layer, in multiples of 128
        users (every GEMM row count stays a multiple of the 256-row CTA-pair tile), split evenly over the call."""
        per_user = S * 2 * self.config.hidden_size * 2
        n = max(1, int(self.kv_chunk_bytes // max(per_user, 1)))
        if n >= B:
            return B
        chunks = -(-B // n)
        n = -(-B // chunks)
        return -(-n // 128) * 128 if n >= 128 else n

    def encode_from_kv(self, query_embeddings: torch.Tensor, kv_all: Optional[torch.Tensor], B: int, S: int,
                       mask: Optional[torch.Tensor], out_dtype: torch.dtype = torch.float32,
                       prelayernorm_dtype: torch.dtype = torch.float32, enc: Optional[torch.Tensor] = None,
                       kv_chunk_users: Optional[int] = None) -> torch.Tensor:
        """The encoder behind the cross-attention K/V projection: kv_all bf16 [B * S, 2 H n_cross] (K and V of every
        cross-attention layer side by side), mask fp32 [B, S] (1 attend / 0 masked) or None.  With `enc` (bf16
        [B * S, E], kv_all = None) the cross-attention layers project K / V themselves: `kv_chunk_users` = None runs the
        fused K/V-projection + attention kernel on it; `kv_chunk_users` = n materialises one layer's K/V for n users at a
        time (GEMM + attention per chunk) - the LAYER-MAJOR order of the user model: the K/V buffer stays bounded while the
        self-attention / output / FFN GEMMs of the layer see all B * Q rows in one launch (at B = 4096 they run at 0.92-1.03
        of the sustained bf16 peak against 0.66-0.85 at the 512-user chunks of a chunk-major loop)."""
        cfg = self.config
        H, heads = cfg.hidden_size, cfg.num_attention_heads
        Q = query_embeddings.shape[1]
        pk = self.packed()
        if (self.fold_layernorm and kv_all is not None and prelayernorm_dtype == torch.bfloat16 and H % 256 == 0 and
                cfg.intermediate_size % 256 == 0 and len(pk["layers"]) > 0):
            return self._encode_from_kv_folded(query_embeddings, kv_all, B, S, mask, out_dtype)

        # BertEmbeddings query-only branch: LayerNorm of the learned tokens - batch-invariant, and so is everything up
        # to the first cross-attention: layer 0's whole self-attention block and its cross-attention QUERY projection
        # see the same [Q, H] rows for every batch element.  They are computed once on Q rows per (weights, queries)
        # version and cached (`_layer0_invariants`); per call only the LayerNorm that ends the block runs on B*Q rows
        # (it broadcasts its Q input rows) and the cross-attention reads one shared set of queries (q_broadcast).
        # Exactly the reference's arithmetic (models/qformer.py:103-108, 417-432) with the B-fold repetition removed:
        # 3.1 % of the item model's FLOPs, 1.9 % of the user model's.
        nl = len(pk["layers"])
        hoist = self.hoist_layer0 and nl > 0 and pk["layers"][0]["cross"]
        if hoist:
            inv = self._layer0_invariants(pk, query_embeddings, prelayernorm_dtype)
        else:
            q0 = query_embeddings.detach().reshape(Q, H).float().contiguous()
            h = ops.layernorm(q0, pk["emb_g"], pk["emb_b"], cfg.layer_norm_eps, rows=B * Q, in_row_mod=Q)

        for li, L in enumerate(pk["layers"]):
            last = li == nl - 1
            if li == 0 and hoist:
                h = ops.layernorm(inv["pre1"], L["ln1_g"], L["ln1_b"], cfg.layer_norm_eps, rows=B * Q, in_row_mod=Q)
            else:
                qkv = ops.linear(h, L["w_qkv"], L["b_qkv"])
                ctx = ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], batch=B, num_heads=heads, nq=Q, nk=Q)
                pre = ops.linear(ctx, L["w_o"], L["b_o"], epilogue=ops.EPI_BIAS_RESIDUAL, residual=h,
                                 out_dtype=prelayernorm_dtype)
                h = ops.layernorm(pre, L["ln1_g"], L["ln1_b"], cfg.layer_norm_eps)
            if L["cross"]:
                off = L["kv_slot"] * 2 * H
                if enc is not None and kv_chunk_users is not None:
                    shared_q = li == 0 and hoist
                    qc = inv["qc"] if shared_q else ops.linear(h, L["w_qc"], L["b_qc"])
                    ctx = torch.empty(B * Q, H, device=enc.device, dtype=torch.bfloat16)
                    w_kv, b_kv = pk["w_kv_all"][off:off + 2 * H], pk["b_kv_all"][off:off + 2 * H]
                    for lo in range(0, B, kv_chunk_users):
                        hi = min(lo + kv_chunk_users, B)
                        kv = ops.linear(enc[lo * S:hi * S], w_kv, b_kv)
                        ops.attention(qc if shared_q else qc[lo * Q:hi * Q], kv[:, :H], kv[:, H:], batch=hi - lo,
                                      num_heads=heads, nq=Q, nk=S, key_mask=None if mask is None else mask[lo:hi],
                                      q_broadcast=shared_q, out=ctx[lo * Q:hi * Q])
                        del kv
                elif enc is not None:
                    if li == 0 and hoist:
                        ctx = ops.kv_attention(enc, L["w_kvp"], inv["qc"], L["b_v"], batch=B, num_heads=heads, nk=S,
                                               key_mask=mask, q_broadcast=True)
                    else:
                        ctx = ops.kv_attention(enc, L["w_kvp"], ops.linear(h, L["w_qc"], L["b_qc"]), L["b_v"], batch=B,
                                               num_heads=heads, nk=S, key_mask=mask)
                elif li == 0 and hoist:
                    ctx = ops.attention(inv["qc"], kv_all[:, off:off + H], kv_all[:, off + H:off + 2 * H], batch=B,
                                        num_heads=heads, nq=Q, nk=S, key_mask=mask, q_broadcast=True)
                else:
                    qc = ops.linear(h, L["w_qc"], L["b_qc"])
                    ctx = ops.attention(qc, kv_all[:, off:off + H], kv_all[:, off + H:off + 2 * H], batch=B,
                                        num_heads=heads, nq=Q, nk=S, key_mask=mask)
                pre = ops.linear(ctx, L["w_oc"], L["b_oc"], epilogue=ops.EPI_BIAS_RESIDUAL, residual=h,
                                 out_dtype=prelayernorm_dtype)
                h = ops.layernorm(pre, L["ln2_g"], L["ln2_b"], cfg.layer_norm_eps)
            inter = ops.linear(h, L["w_1"], L["b_1"], epilogue=ops.EPI_BIAS_GELU)
            pre = ops.linear(inter, L["w_2"], L["b_2"], epilogue=ops.EPI_BIAS_RESIDUAL, residual=h,
                             out_dtype=prelayernorm_dtype)
            h = ops.layernorm(pre, L["ln3_g"], L["ln3_b"], cfg.layer_norm_eps,
                              out_dtype=out_dtype if last else torch.bfloat16)
        return h.view(B, Q, H)


    def _folded_pack(self, pk: dict) -> list:
        """Per layer, the weights of the GEMMs that read a LayerNorm output, with that LayerNorm folded in
        (ops.fold_layernorm_weights): qkv <- the previous layer's closing LayerNorm, the cross-attention query projection <-
        ln1, FFN-up <- ln2 (cross layers) or ln1.  Cached with the weight pack."""
        fp = pk.get("folded")
        if fp is None:
            fp, prev = [], None
            for L, lyr in zip(pk["layers"], self.encoder.layer):
                a, d = getattr(lyr.attention, "self"), {}
                if prev is not None:
                    # folded from the fp32 master weights (one bf16 rounding of W o gamma)
                    w = torch.cat([a.query.weight, a.key.weight, a.value.weight], 0)
                    d["qkv"] = ops.fold_layernorm_weights(w, L["b_qkv"], prev["ln3_g"], prev["ln3_b"])
                w1 = lyr.intermediate_query.dense.weight
                if L["cross"]:
                    c = getattr(lyr.crossattention, "self")
                    d["qc"] = ops.fold_layernorm_weights(c.query.weight, L["b_qc"], L["ln1_g"], L["ln1_b"])
                    d["w1"] = ops.fold_layernorm_weights(w1, L["b_1"], L["ln2_g"], L["ln2_b"])
                else:
                    d["w1"] = ops.fold_layernorm_weights(w1, L["b_1"], L["ln1_g"], L["ln1_b"])
                fp.append(d)
                prev = L
            pk["folded"] = fp
        return fp

    def _encode_from_kv_folded(self, query_embeddings: torch.Tensor, kv_all: torch.Tensor, B: int, S: int,
                               mask: Optional[torch.Tensor], out_dtype: torch.dtype) -> torch.Tensor:
        """`encode_from_kv` without LayerNorm kernels (models/qformer.py:285-289, 371-375: h = LayerNorm(dense(x) + input)).
        A hidden state lives as `pre` (bf16, the LayerNorm INPUT) + its fp32 row statistics, written by the residual GEMM
        that produced it; its consumers apply the normalisation themselves: as an A operand through gamma-scaled weights and
        a rank-one correction in the epilogue, as a residual through (x - mu) rstd gamma + beta on the residual tile.  Only
        the batch-broadcast LayerNorm of the hoisted layer 0 and the LayerNorm that closes the encoder are still kernels:
        per layer 2-3 fewer streaming passes over [B Q, H] (read + write each)."""
        cfg = self.config
        H, heads, eps = cfg.hidden_size, cfg.num_attention_heads, cfg.layer_norm_eps
        Q = query_embeddings.shape[1]
        pk = self.packed()
        fp = self._folded_pack(pk)
        nl = len(pk["layers"])
        M = B * Q
        hoist = self.hoist_layer0 and pk["layers"][0]["cross"]
        stats = ops.ln_stats_buffer(M, H, kv_all.device, count=3 * nl)
        RES = ops.EPI_BIAS_RESIDUAL
        # the current hidden state: either `h` (a materialised LayerNorm output) or (`pre`, `st`, gamma, beta)
        h = pre = st = g = b = None
        if hoist:
            inv = self._layer0_invariants(pk, query_embeddings, torch.bfloat16)
        else:
            q0 = query_embeddings.detach().reshape(Q, H).float().contiguous()
            h = ops.layernorm(q0, pk["emb_g"], pk["emb_b"], eps, rows=M, in_row_mod=Q)

        def dense_residual(x, w, bias, slot):
            # pre' = dense(x) + hidden, statistics of pre' into stats[slot]
            if h is not None:
                return ops.linear_ln(x, w, bias, epilogue=RES, residual=h, stats_out=stats[slot], eps=eps, hidden=H)
            return ops.linear_ln(x, w, bias, epilogue=RES, residual=pre, ln_res=(st, g, b), stats_out=stats[slot],
                                 eps=eps, hidden=H)

        for li, (L, F) in enumerate(zip(pk["layers"], fp)):
            if li == 0 and hoist:
                h = ops.layernorm(inv["pre1"], L["ln1_g"], L["ln1_b"], eps, rows=M, in_row_mod=Q)
            else:
                if h is not None:
                    qkv = ops.linear(h, L["w_qkv"], L["b_qkv"])
                else:
                    qkv = ops.linear_ln(pre, F["qkv"][0], F["qkv"][1], ln_in=(st, F["qkv"][2]), eps=eps, hidden=H)
                ctx = ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], batch=B, num_heads=heads, nq=Q, nk=Q)
                pre_n = dense_residual(ctx, L["w_o"], L["b_o"], 3 * li)
                h, pre, st, g, b = None, pre_n, stats[3 * li], L["ln1_g"], L["ln1_b"]
            if L["cross"]:
                off = L["kv_slot"] * 2 * H
                if li == 0 and hoist:
                    ctx = ops.attention(inv["qc"], kv_all[:, off:off + H], kv_all[:, off + H:off + 2 * H], batch=B,
                                        num_heads=heads, nq=Q, nk=S, key_mask=mask, q_broadcast=True)
                else:
                    if h is not None:
                        qc = ops.linear(h, L["w_qc"], L["b_qc"])
                    else:
                        qc = ops.linear_ln(pre, F["qc"][0], F["qc"][1], ln_in=(st, F["qc"][2]), eps=eps, hidden=H)
                    ctx = ops.attention(qc, kv_all[:, off:off + H], kv_all[:, off + H:off + 2 * H], batch=B,
                                        num_heads=heads, nq=Q, nk=S, key_mask=mask)
                pre_n = dense_residual(ctx, L["w_oc"], L["b_oc"], 3 * li + 1)
                h, pre, st, g, b = None, pre_n, stats[3 * li + 1], L["ln2_g"], L["ln2_b"]
            if h is not None:
                inter = ops.linear(h, L["w_1"], L["b_1"], epilogue=ops.EPI_BIAS_GELU)
            else:
                inter = ops.linear_ln(pre, F["w1"][0], F["w1"][1], epilogue=ops.EPI_BIAS_GELU, ln_in=(st, F["w1"][2]),
                                      eps=eps, hidden=H)
            pre_n = dense_residual(inter, L["w_2"], L["b_2"], 3 * li + 2)
            h, pre, st, g, b = None, pre_n, stats[3 * li + 2], L["ln3_g"], L["ln3_b"]
        out = ops.layernorm(pre, g, b, eps, out_dtype=out_dtype)
        return out.view(B, Q, H)


def _check_inference_mode(module: nn.Module, dropout: float):
    if module.training and dropout > 0.0:
        raise NotImplementedError(
            "unirec_b200: this entry point implements dropout = identity; call .eval() first.  Train-mode dropout is "
            "implemented for QFormerForItemRepresentation.forward (unirec_b200/training.py).")


class QFormerForItemRepresentation(nn.Module):
    """Item Q-Former (reference: models/qformer_utils.py:16-60)."""

    def __init__(self, hidden_size: int = 1024, num_hidden_layers: int = 12, num_attention_heads: int = 16,
                 intermediate_size: int = 4096, num_query_tokens: int = 32, field_embedding_dim: int = 1024,
                 num_fields: int = None, dropout: float = 0.2):
        super().__init__()
        if num_fields is None:
            raise ValueError("num_fields must be provided")
        self.config = _bert_config(
            hidden_size=hidden_size, num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
            intermediate_size=intermediate_size, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout,
            add_cross_attention=True, query_length=num_query_tokens, encoder_width=field_embedding_dim,
            cross_attention_freq=2)
        self.num_query_tokens = num_query_tokens
        self.query_embeddings = nn.Parameter(torch.randn(1, num_query_tokens, hidden_size))
        self.qformer = QFormerBackbone(self.config)
        self.item_representation_head = nn.Linear(hidden_size, field_embedding_dim)
        self.reconstruction_head = nn.Linear(hidden_size, field_embedding_dim)
        self.field_projection = nn.Linear(num_query_tokens, num_fields)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())
        self.output_dtype = torch.float32          # the reference returns fp32 tensors
        self.prelayernorm_dtype = torch.float32
        self._head_pack = None
        self._head_key = None
        # train-mode dropout: None = draw a fresh seed per forward call from torch's default CPU generator
        # (reproducible under torch.manual_seed); an int = use it for the next call, then count up
        self.dropout_seed: Optional[int] = None
        self.last_dropout = None                   # (thr16, seed) of the most recent train-mode forward
        # optional int64 CUDA tensor with one element, added to the seed by the kernels at run time: a training step
        # captured in a CUDA graph increments it inside the graph (training.TrainStepGraph)
        self.dropout_seed_offset: Optional[torch.Tensor] = None

    def _next_dropout(self):
        p = float(self.config.hidden_dropout_prob)
        if p != float(self.config.attention_probs_dropout_prob):
            raise NotImplementedError("hidden and attention dropout probabilities differ")
        if p <= 0.0:
            self.last_dropout = None
            return None
        if self.dropout_seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        else:
            seed = int(self.dropout_seed)
            self.dropout_seed = seed + 1
        self.last_dropout = (ops.dropout_threshold(p), seed)
        if self.dropout_seed_offset is not None:
            return self.last_dropout + (self.dropout_seed_offset,)
        return self.last_dropout

    def invalidate_packed(self):
        """Forget every cached bf16 weight copy (backbone and heads); see QFormerBackbone.invalidate_packed."""
        self._head_pack = self._head_key = None
        self.qformer.invalidate_packed()

    def _apply(self, fn, *a, **k):
        self._head_pack = self._head_key = None
        return super()._apply(fn, *a, **k)

    def _heads(self):
        ps = [self.item_representation_head.weight, self.item_representation_head.bias,
              self.reconstruction_head.weight, self.reconstruction_head.bias,
              self.field_projection.weight, self.field_projection.bias]
        key = (ps[0].device, tuple(p._version for p in ps), ps[0].data_ptr())
        if self._head_pack is None or self._head_key != key:
            self._head_pack = {
                "w_rep": ps[0].detach().to(torch.bfloat16).contiguous(), "b_rep": ps[1].detach().float().contiguous(),
                "w_rec": ps[2].detach().to(torch.bfloat16).contiguous(), "b_rec": ps[3].detach().float().contiguous(),
                "w_fp": ps[4].detach().float().contiguous(), "b_fp": ps[5].detach().float().contiguous(),
            }
            self._head_key = key
        return self._head_pack

    @torch.no_grad()
    def encode_query_tokens(self, field_embeddings: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                            out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        """query_outputs only ([B, Q, H]); what batched item-token generation needs
        (data_processing/qformer_inference.py:160-163)."""
        _check_inference_mode(self, self.config.hidden_dropout_prob)
        return self.qformer.encode(self.query_embeddings, field_embeddings, attention_mask, out_dtype,
                                   self.prelayernorm_dtype)

    def _forward_train(self, field_embeddings, attention_mask, drop=None):
        """Differentiable forward (models/qformer_utils.py:37-60 under autograd): backbone = BackboneTrainFn,
        heads = LinearFn (tcgen05 fwd / dgrad / wgrad), the 32 -> num_fields projection in torch (0.9 GFLOP/1024 items).
        drop = (thr16, seed[, seed_offset]) or None."""
        from .training import BackboneTrainFn, LinearFn
        bb = self.qformer
        qo = BackboneTrainFn.apply(bb, drop, field_embeddings, attention_mask, self.query_embeddings, *bb._live_params())
        B, Q, H = qo.shape
        qo16 = qo.to(torch.bfloat16)
        rep = LinearFn.apply(qo16.mean(dim=1).to(torch.bfloat16), self.item_representation_head.weight,
                             self.item_representation_head.bias).float()
        rec = LinearFn.apply(qo16.reshape(B * Q, H), self.reconstruction_head.weight,
                             self.reconstruction_head.bias).view(B, Q, -1).float()
        fields = torch.nn.functional.linear(rec.transpose(1, 2), self.field_projection.weight,
                                            self.field_projection.bias).transpose(1, 2)
        return {"query_outputs": qo, "item_representation": rep, "reconstructed_fields": fields}

    def forward(self, field_embeddings: torch.Tensor, attention_mask: torch.Tensor = None) -> Dict[str, torch.Tensor]:
        if self.training:
            # train(): dropout is live in every forward, including the no-grad positive / negative passes of the
            # reference's step (training/item_qformer_training.py:122-125)
            drop = self._next_dropout()
            if torch.is_grad_enabled() or drop is not None:
                return self._forward_train(field_embeddings, attention_mask, drop)
        with torch.no_grad():
            return self._forward_eval(field_embeddings, attention_mask)

    def _forward_eval(self, field_embeddings: torch.Tensor, attention_mask: torch.Tensor = None) -> Dict[str, torch.Tensor]:
        od = self.output_dtype
        query_outputs = self.qformer.encode(self.query_embeddings, field_embeddings, attention_mask, od,
                                            self.prelayernorm_dtype)
        hp = self._heads()
        qo = ops.cast_bf16(query_outputs)
        item_representation = ops.linear(ops.mean_tokens(qo), hp["w_rep"], hp["b_rep"], out_dtype=od)
        rec = ops.linear(qo, hp["w_rec"], hp["b_rec"])
        reconstructed_fields = ops.field_projection(rec, hp["w_fp"], hp["b_fp"], out_dtype=od)
        return {"query_outputs": query_outputs, "item_representation": item_representation,
                "reconstructed_fields": reconstructed_fields}


class UserQFormer(nn.Module):
    """User Q-Former (reference: training/user_qformer_training.py:17-68)."""

    def __init__(self, hidden_size: int = 1024, num_hidden_layers: int = 4, num_attention_heads: int = 16,
                 intermediate_size: int = 4096, num_query_tokens: int = 64, input_embedding_dim: int = 1024,
                 num_item_tokens_to_predict: int = 32, dropout: float = 0.1):
        super().__init__()
        self.config = _bert_config(
            hidden_size=hidden_size, num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
            intermediate_size=intermediate_size, hidden_dropout_prob=dropout, attention_probs_dropout_prob=dropout,
            add_cross_attention=True, query_length=num_query_tokens, encoder_width=input_embedding_dim,
            cross_attention_freq=1)
        self.num_query_tokens = num_query_tokens
        self.query_embeddings = nn.Parameter(torch.randn(1, num_query_tokens, hidden_size))
        self.qformer = QFormerBackbone(self.config)
        self.prediction_head = nn.Sequential(
            nn.Linear(hidden_size, hidden_size),
            nn.GELU(),
            nn.LayerNorm(hidden_size),
            nn.Linear(hidden_size, num_item_tokens_to_predict * input_embedding_dim))
        self.num_item_tokens_to_predict = num_item_tokens_to_predict
        self.input_embedding_dim = input_embedding_dim
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())
        self.output_dtype = torch.float32
        self.prelayernorm_dtype = torch.float32
        # Chunk-major (default): the cross-attention K/V of ALL layers are materialised per chunk of users by one GEMM that
        # reads the user sequence once: chunk * S * layers * 2 * H * 2 bytes (13.4 GB for 512 users x 1600 keys x 4 layers)
        self.max_kv_bytes = 14 << 30
        # Layer-major (`layer_major = True`): one encoder call covers up to max_seq_bytes of user sequence (4096 users at
        # S = 1600) and materialises the K/V of ONE layer per chunk of users inside it (QFormerBackbone.encode) - bounded
        # K/V memory for any number of layers, but the sequence is re-read once per layer: measured 23.4 k users/s against
        # 25.3 k chunk-major on one box (profiles/r02_c_*: DRAM traffic of the K/V GEMM 26.3 GB instead of 17 GB per
        # 13.7 TFLOP, and on a power-capped part those bytes are SM clock, 1215 vs 1413 MHz) - so it is off by default
        self.layer_major = False
        self.max_seq_bytes = 14 << 30
        self._head_pack = None
        self._head_key = None

    def invalidate_packed(self):
        """Forget every cached bf16 weight copy (backbone and head); see QFormerBackbone.invalidate_packed."""
        self._head_pack = self._head_key = None
        self.qformer.invalidate_packed()

    def _apply(self, fn, *a, **k):
        self._head_pack = self._head_key = None
        return super()._apply(fn, *a, **k)

    def _heads(self):
        ph = self.prediction_head
        ps = [ph[0].weight, ph[0].bias, ph[2].weight, ph[2].bias, ph[3].weight, ph[3].bias]
        key = (ps[0].device, tuple(p._version for p in ps), ps[0].data_ptr())
        if self._head_pack is None or self._head_key != key:
            self._head_pack = {
                "w0": ps[0].detach().to(torch.bfloat16).contiguous(), "b0": ps[1].detach().float().contiguous(),
                "g": ps[2].detach().float().contiguous(), "b": ps[3].detach().float().contiguous(),
                "w3": ps[4].detach().to(torch.bfloat16).contiguous(), "b3": ps[5].detach().float().contiguous(),
            }
            self._head_key = key
        return self._head_pack

    @property
    def fused_kv_attention(self) -> bool:
        return self.qformer.fused_kv_attention

    @fused_kv_attention.setter
    def fused_kv_attention(self, on: bool):
        self.qformer.fused_kv_attention = bool(on)

    def _chunk_users(self, S: int) -> int:
        """Users per encoder call, in multiples of 128 users (every GEMM row count stays a multiple of the 256-row
        CTA-pair tile and the tile counts of the small per-chunk GEMMs stay close to whole waves of 74 clusters)."""
        cfg = self.config
        if self.qformer.fused_kv_supported(S) or self.layer_major:
            # no K/V buffer for all layers: a call is bounded by the bf16 user sequence itself
            self.qformer.kv_chunk_bytes = self.max_kv_bytes if self.layer_major else None
            n = max(1, int(self.max_seq_bytes // max(S * cfg.encoder_width * 2, 1)))
        else:
            self.qformer.kv_chunk_bytes = None
            n = max(1, int(self.max_kv_bytes // max(S * cfg.num_hidden_layers * 2 * cfg.hidden_size * 2, 1)))
        return (n // 128) * 128 if n >= 128 else n

    @torch.no_grad()
    def encode_queries(self, user_sequence_tokens: torch.Tensor, attention_mask: Optional[torch.Tensor],
                       out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        """last_hidden_state [B, num_query_tokens, H] of the user Q-Former."""
        _check_inference_mode(self, self.config.hidden_dropout_prob)
        B, S, _ = user_sequence_tokens.shape
        step = self._chunk_users(S)
        outs = []
        for lo in range(0, B, step):
            m = None if attention_mask is None else attention_mask[lo:lo + step]
            outs.append(self.qformer.encode(self.query_embeddings, user_sequence_tokens[lo:lo + step], m, out_dtype,
                                            self.prelayernorm_dtype))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def _position_tables(self, hmax: int, tokens_per_item: int, device) -> dict:
        """Per (history length, weight version): posbias = PE W_kv^T (bf16 [S + 128, 2 H layers], the first 128 rows
        repeated at the end) and pad = -PE (bf16 [S, E]) for `ops.linear_gather`.  PE is split into two bf16 terms so that
        posbias carries fp32-accurate positions; it is rounded to bf16 once."""
        pk = self.qformer.packed()
        key = ("pos", hmax, tokens_per_item)
        t = pk.get(key)
        if t is None:
            S, E = hmax * tokens_per_item, self.config.encoder_width
            pe = ops.positional_encoding_table(S, E, device)
            hi = pe.to(torch.bfloat16)
            lo = (pe - hi.float()).to(torch.bfloat16)
            w = pk["w_kv_all"]
            posw = (ops.linear(hi, w, None, out_dtype=torch.float32) + ops.linear(lo, w, None, out_dtype=torch.float32))
            posw = posw.to(torch.bfloat16)
            t = {"posbias": torch.cat([posw, posw[:128]], 0).contiguous(), "pad": (-hi).contiguous()}
            pk[key] = t
        return t

    @torch.no_grad()
    def encode_queries_from_history(self, item_tokens: torch.Tensor, history: torch.Tensor, lengths: torch.Tensor,
                                    out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        """last_hidden_state [B, num_query_tokens, H] straight from the item-token table: the user
        sequence (gather + sinusoidal PE + right padding, models/user_sequence_encoder.py:128-140,
        training/user_qformer_training.py:153-161) is never built - the K/V projection gathers its A operand from the
        table and adds the position term as a precomputed tile (`ops.linear_gather`).  item_tokens bf16 [N, 32, E],
        history int64 [B, Hmax], lengths int32 [B].  No per-user context vector on this path (the builder path has it)."""
        _check_inference_mode(self, self.config.hidden_dropout_prob)
        if item_tokens.dim() != 3 or item_tokens.shape[1] != 32:
            raise RuntimeError("encode_queries_from_history: needs an item-token table [N, 32, E]")
        B, hmax = history.shape
        Q_item = item_tokens.shape[1]
        S = hmax * Q_item
        pk = self.qformer.packed()
        tabs = self._position_tables(hmax, Q_item, item_tokens.device)
        # this path materialises the K/V of ALL layers per chunk of users (one gathered GEMM): 512 users at S = 1600
        cfg = self.config
        step = max(1, int(self.max_kv_bytes // max(S * cfg.num_hidden_layers * 2 * cfg.hidden_size * 2, 1)))
        step = (step // 128) * 128 if step >= 128 else step
        self.qformer.kv_chunk_bytes = None
        pos = torch.arange(S, device=item_tokens.device, dtype=torch.int32)
        outs = []
        for lo in range(0, B, step):
            hi = min(lo + step, B)
            hist, lens = history[lo:hi].contiguous(), lengths[lo:hi].contiguous()
            kv_all = ops.linear_gather(item_tokens, hist, lens, tabs["pad"], pk["w_kv_all"], pk["b_kv_all"], tabs["posbias"])
            mask = (pos.unsqueeze(0) < (lens * Q_item).unsqueeze(1)).to(torch.float32)
            outs.append(self.qformer.encode_from_kv(self.query_embeddings, kv_all, hi - lo, S, mask, out_dtype,
                                                    self.prelayernorm_dtype))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    @torch.no_grad()
    def predict_from_queries(self, hidden: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """mean-pool + prediction head (training/user_qformer_training.py:60-66)."""
        hp = self._heads()
        od = self.output_dtype if out_dtype is None else out_dtype
        B = hidden.shape[0]
        rep = ops.mean_tokens(ops.cast_bf16(hidden))
        g = ops.linear(rep, hp["w0"], hp["b0"], epilogue=ops.EPI_BIAS_GELU)
        g = ops.layernorm(g, hp["g"], hp["b"], self.prediction_head[2].eps)
        flat = ops.linear(g, hp["w3"], hp["b3"], out_dtype=od)
        return flat.view(B, self.num_item_tokens_to_predict, self.input_embedding_dim)

    @torch.no_grad()
    def forward(self, user_sequence_tokens: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        hidden = self.encode_queries(user_sequence_tokens, attention_mask, torch.bfloat16)
        return self.predict_from_queries(hidden)
