"""Deterministic synthetic weights and inputs of the reference shapes.

The reference ships no checkpoints and no data, so every
benchmark and parity test runs on "random-init weights of the reference shapes"
(reference init rule: models/qformer.py:664-674, N(0, 0.02) for Linear/Embedding
weights; query tokens ~ randn, models/qformer_utils.py:30).  To make the SAME
weights reproducible in three places (the reference itself when pinning the
oracle, the CPU oracle, the CUDA modules) without shipping a 1.2 GB checkpoint,
every tensor is generated from numpy's PCG64 seeded by crc32(key) ^ seed: the
value of a tensor depends only on (key, shape, seed), never on construction order.
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import numpy as np
import torch


def _rng(key: str, seed: int) -> np.random.Generator:
    return np.random.default_rng((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF)


def normal(key: str, shape, seed: int = 0, std: float = 1.0, mean: float = 0.0) -> torch.Tensor:
    a = _rng(key, seed).standard_normal(size=tuple(shape), dtype=np.float32)
    if std != 1.0:
        a *= np.float32(std)
    if mean != 0.0:
        a += np.float32(mean)
    return torch.from_numpy(a)


def qformer_backbone_shapes(hidden: int, layers: int, inter: int, enc_width: int,
                            cross_freq: int, vocab: int = 30522, max_pos: int = 512
                            ) -> Dict[str, Tuple[int, ...]]:
    """Shapes of every tensor in the reference BertModel state dict (models/qformer.py:51-76,
    111-134, 278-283, 349-368, 378-400), INCLUDING the tensors the query-only path never
    executes (word/position embeddings, text-branch FFN) - they must round-trip through
    checkpoints."""
    s: Dict[str, Tuple[int, ...]] = {
        "embeddings.word_embeddings.weight": (vocab, hidden),
        "embeddings.position_embeddings.weight": (max_pos, hidden),
        "embeddings.LayerNorm.weight": (hidden,),
        "embeddings.LayerNorm.bias": (hidden,),
    }
    for i in range(layers):
        p = f"encoder.layer.{i}."
        blocks = [("attention", hidden)]
        if i % cross_freq == 0:
            blocks.append(("crossattention", enc_width))
        for name, kv_in in blocks:
            s[p + f"{name}.self.query.weight"] = (hidden, hidden)
            s[p + f"{name}.self.query.bias"] = (hidden,)
            s[p + f"{name}.self.key.weight"] = (hidden, kv_in)
            s[p + f"{name}.self.key.bias"] = (hidden,)
            s[p + f"{name}.self.value.weight"] = (hidden, kv_in)
            s[p + f"{name}.self.value.bias"] = (hidden,)
            s[p + f"{name}.output.dense.weight"] = (hidden, hidden)
            s[p + f"{name}.output.dense.bias"] = (hidden,)
            s[p + f"{name}.output.LayerNorm.weight"] = (hidden,)
            s[p + f"{name}.output.LayerNorm.bias"] = (hidden,)
        for suffix in ("", "_query"):
            s[p + f"intermediate{suffix}.dense.weight"] = (inter, hidden)
            s[p + f"intermediate{suffix}.dense.bias"] = (inter,)
            s[p + f"output{suffix}.dense.weight"] = (hidden, inter)
            s[p + f"output{suffix}.dense.bias"] = (hidden,)
            s[p + f"output{suffix}.LayerNorm.weight"] = (hidden,)
            s[p + f"output{suffix}.LayerNorm.bias"] = (hidden,)
    return s


def _fill(shapes: Dict[str, Tuple[int, ...]], seed: int, attn_std: float, live_only: bool
          ) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    for k, shp in shapes.items():
        dead = ("word_embeddings" in k or "position_embeddings" in k
                or ".intermediate.dense" in k or ".output.dense" in k or ".output.LayerNorm" in k)
        # ".output.dense" above also matches "attention.output.dense"; exclude those from "dead".
        if "attention.output" in k:
            dead = False
        if dead and live_only:
            out[k] = torch.zeros(shp)
            continue
        if k.endswith("LayerNorm.weight") or k.endswith("prediction_head.2.weight"):
            out[k] = normal(k, shp, seed, std=0.1, mean=1.0)
        elif k.endswith("LayerNorm.bias") or k.endswith("prediction_head.2.bias"):
            out[k] = normal(k, shp, seed, std=0.1)
        elif k.endswith(".bias"):
            out[k] = normal(k, shp, seed, std=0.02)
        elif k.endswith("query_embeddings"):
            out[k] = normal(k, shp, seed, std=1.0)
        elif ".self.query.weight" in k or ".self.key.weight" in k:
            out[k] = normal(k, shp, seed, std=attn_std)
        else:
            out[k] = normal(k, shp, seed, std=0.02)
    return out


def item_qformer_state_dict(hidden=1024, layers=12, inter=4096, num_query=32, field_dim=1024,
                            num_fields=14, seed=0, attn_std=0.02, live_only=False,
                            vocab=30522, max_pos=512) -> Dict[str, torch.Tensor]:
    """State dict with the keys of QFormerForItemRepresentation (models/qformer_utils.py:16-35)."""
    shapes = {"query_embeddings": (1, num_query, hidden)}
    for k, v in qformer_backbone_shapes(hidden, layers, inter, field_dim, 2, vocab, max_pos).items():
        shapes["qformer." + k] = v
    shapes.update({
        "item_representation_head.weight": (field_dim, hidden),
        "item_representation_head.bias": (field_dim,),
        "reconstruction_head.weight": (field_dim, hidden),
        "reconstruction_head.bias": (field_dim,),
        "field_projection.weight": (num_fields, num_query),
        "field_projection.bias": (num_fields,),
    })
    sd = _fill(shapes, seed, attn_std, live_only)
    sd["qformer.embeddings.position_ids"] = torch.arange(max_pos).expand((1, -1)).clone()
    return sd


def user_qformer_state_dict(hidden=1024, layers=4, inter=4096, num_query=64, input_dim=1024,
                            num_predict=32, seed=0, attn_std=0.02, live_only=False,
                            vocab=30522, max_pos=512) -> Dict[str, torch.Tensor]:
    """State dict with the keys of UserQFormer (training/user_qformer_training.py:21-45)."""
    shapes = {"query_embeddings": (1, num_query, hidden)}
    for k, v in qformer_backbone_shapes(hidden, layers, inter, input_dim, 1, vocab, max_pos).items():
        shapes["qformer." + k] = v
    shapes.update({
        "prediction_head.0.weight": (hidden, hidden),
        "prediction_head.0.bias": (hidden,),
        "prediction_head.2.weight": (hidden,),
        "prediction_head.2.bias": (hidden,),
        "prediction_head.3.weight": (num_predict * input_dim, hidden),
        "prediction_head.3.bias": (num_predict * input_dim,),
    })
    sd = _fill(shapes, seed, attn_std, live_only)
    sd["qformer.embeddings.position_ids"] = torch.arange(max_pos).expand((1, -1)).clone()
    return sd


def item_fields(batch: int, num_fields: int = 14, dim: int = 1024, seed: int = 1,
                clip_field: int = 7, presence: float = 1.0, all_masked_row: int = -1):
    """Synthetic field embeddings (cfg 1): randn, the CLIP image field
    zero-padded from col 768 (models/item_encoder_pure_value.py:163,257 pads 768-d CLIP to
    1024), optional Bernoulli field presence (absent field = zero vector = mask 0,
    models/qformer_utils.py:116) and optionally one row with every field masked."""
    x = normal("item_fields", (batch, num_fields, dim), seed)
    if dim > 768 and 0 <= clip_field < num_fields:
        x[:, clip_field, (dim * 3) // 4:] = 0
    mask = torch.ones(batch, num_fields, dtype=torch.long)
    if presence < 1.0:
        keep = _rng("item_presence", seed).random((batch, num_fields)) < presence
        keep[:, 0] = True
        mask = torch.from_numpy(keep.astype(np.int64))
    if 0 <= all_masked_row < batch:
        mask[all_masked_row] = 0
    x = x * mask[..., None].to(x.dtype)
    return x, mask


def user_sequences(batch: int, max_items: int = 50, tokens_per_item: int = 32, dim: int = 1024,
                   seed: int = 2, ragged: bool = True):
    """Synthetic padded user sequences [B, S, dim] + mask [B, S] float (collate_fn contract,
    training/user_qformer_training.py:153-161): right-padded with zeros, lengths multiples of
    tokens_per_item."""
    S = max_items * tokens_per_item
    x = normal("user_seq", (batch, S, dim), seed)
    if ragged:
        n_items = _rng("user_len", seed).integers(1, max_items + 1, size=batch)
        n_items[0] = max_items
    else:
        n_items = np.full(batch, max_items)
    lens = torch.from_numpy(n_items * tokens_per_item)
    mask = (torch.arange(S)[None, :] < lens[:, None]).float()
    x = x * mask[..., None]
    return x, mask
