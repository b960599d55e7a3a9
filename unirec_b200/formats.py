"""On-disk formats either side of item-token generation.

Reference formats, read and written bit-compatibly (a file written here loads in the reference's code and vice versa):
  * field-embedding cache directory (models/qformer_utils.py:121-147, QFormerDataset._load_cache/_save_cache):
      embeddings.pt  torch.save({sample_idx: FloatTensor[F, E]})
      masks.pt       torch.save({sample_idx: LongTensor[F]})
      fields.json    json list of the F field names, in sorted order; the cache is valid only if it equals the
                     caller's field list (:129-137)
  * item query-token pickle (data_processing/qformer_inference.py:163-173): pickle.dump({item_id: float32 ndarray[Q, H]})

Those dict-of-small-tensors files are what makes the reference's generation step slow (one Python object per item).
The native table below is what the kernels consume: one contiguous bf16 matrix per shard that can be memory-mapped and
copied to HBM in large pinned chunks, item-range sharded exactly like generation (pipeline.shard_range):
  * token table directory:
      table.json           {"format": "unirec_b200.item_tokens", "version": 1, "num_items", "tokens_per_item", "hidden",
                            "dtype": "bfloat16", "shards": [{"file", "first_item", "num_items"}]}
      tokens.<k>.bf16      raw little-endian bf16 [num_items_k, Q, H], row-major
      item_ids.json        list of item ids in row order (optional)
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

TABLE_FORMAT = "unirec_b200.item_tokens"
TABLE_VERSION = 1


# --------------------------------------------------------------------------------------- field-embedding cache
def save_field_cache(cache_dir: str, field_embeddings: torch.Tensor, attention_mask: torch.Tensor,
                     field_names: Sequence[str]) -> None:
    """Write [N, F, E] embeddings + [N, F] masks as the reference's cache (models/qformer_utils.py:139-145)."""
    if field_embeddings.dim() != 3 or attention_mask.shape != field_embeddings.shape[:2]:
        raise ValueError("expected field_embeddings [N, F, E] and attention_mask [N, F]")
    if len(field_names) != field_embeddings.shape[1]:
        raise ValueError("len(field_names) must equal the number of fields")
    os.makedirs(cache_dir, exist_ok=True)
    emb = field_embeddings.detach().to("cpu", torch.float32)
    msk = attention_mask.detach().to("cpu", torch.long)
    # clone(): a view would drag the whole [N, F, E] storage into every pickled entry
    torch.save({i: emb[i].clone() for i in range(emb.shape[0])}, os.path.join(cache_dir, "embeddings.pt"))
    torch.save({i: msk[i].clone() for i in range(msk.shape[0])}, os.path.join(cache_dir, "masks.pt"))
    with open(os.path.join(cache_dir, "fields.json"), "w") as f:
        json.dump(list(field_names), f)


def load_field_cache(cache_dir: str, expected_fields: Optional[Sequence[str]] = None, pin_memory: bool = False
                     ) -> Optional[Tuple[torch.Tensor, torch.Tensor, List[str]]]:
    """Read the reference's cache into dense tensors ([N, F, E] fp32, [N, F] long, field names), samples in index
    order.  Returns None when a file is missing or the cached field list differs from `expected_fields` - the
    reference's "cache is outdated" rule (models/qformer_utils.py:126-137)."""
    paths = [os.path.join(cache_dir, n) for n in ("embeddings.pt", "masks.pt", "fields.json")]
    if not all(os.path.exists(p) for p in paths):
        return None
    with open(paths[2]) as f:
        fields = json.load(f)
    if expected_fields is not None and list(expected_fields) != fields:
        return None
    emb: Dict[int, torch.Tensor] = torch.load(paths[0], weights_only=False)
    msk: Dict[int, torch.Tensor] = torch.load(paths[1], weights_only=False)
    keys = sorted(emb.keys())
    if keys != sorted(msk.keys()):
        raise ValueError("embeddings.pt and masks.pt hold different sample indices")
    n = len(keys)
    if n == 0:
        return torch.empty(0, len(fields), 0), torch.empty(0, len(fields), dtype=torch.long), fields
    f_, e_ = emb[keys[0]].shape
    out = torch.empty(n, f_, e_, dtype=torch.float32, pin_memory=pin_memory)
    mask = torch.empty(n, f_, dtype=torch.long, pin_memory=pin_memory)
    for r, k in enumerate(keys):
        out[r].copy_(emb[k])
        mask[r].copy_(msk[k])
    return out, mask, fields


# --------------------------------------------------------------------------------------- item-token pickle
def save_item_tokens_pickle(path: str, item_ids: Sequence, tokens: torch.Tensor) -> None:
    """{item_id: float32 ndarray [Q, H]} exactly as data_processing/qformer_inference.py:163-173 writes it."""
    if tokens.dim() != 3 or tokens.shape[0] != len(item_ids):
        raise ValueError("expected tokens [N, Q, H] and N item ids")
    arr = tokens.detach().to("cpu", torch.float32).numpy()
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump({item_id: arr[i].copy() for i, item_id in enumerate(item_ids)}, f)


def load_item_tokens_pickle(path: str, dtype: torch.dtype = torch.bfloat16) -> Tuple[List, torch.Tensor]:
    """Read the reference's pickle into (item ids in file order, dense [N, Q, H] table)."""
    with open(path, "rb") as f:
        d = pickle.load(f)
    ids = list(d.keys())
    if not ids:
        return ids, torch.empty(0, 0, 0, dtype=dtype)
    q, h = np.asarray(d[ids[0]]).shape
    out = torch.empty(len(ids), q, h, dtype=dtype)
    for r, k in enumerate(ids):
        out[r].copy_(torch.from_numpy(np.ascontiguousarray(d[k], dtype=np.float32)))
    return ids, out


# --------------------------------------------------------------------------------------- native token table
def _bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().to("cpu", torch.bfloat16).contiguous().view(torch.int16).numpy().view(np.uint16)


class TokenTableWriter:
    """Streams item-token shards to disk as they come out of generate_item_tokens (one shard per call or rank)."""

    def __init__(self, directory: str, tokens_per_item: int, hidden: int):
        self.dir, self.q, self.h = directory, int(tokens_per_item), int(hidden)
        os.makedirs(directory, exist_ok=True)
        self.shards: List[dict] = []

    def write_shard(self, first_item: int, tokens: torch.Tensor) -> str:
        if tokens.dim() != 3 or tuple(tokens.shape[1:]) != (self.q, self.h):
            raise ValueError(f"expected tokens [n, {self.q}, {self.h}]")
        name = f"tokens.{first_item:012d}.bf16"
        _bf16_bits(tokens).tofile(os.path.join(self.dir, name))
        self.shards.append({"file": name, "first_item": int(first_item), "num_items": int(tokens.shape[0])})
        return name

    def close(self, item_ids: Optional[Sequence] = None, extra_shards: Iterable[dict] = ()) -> None:
        shards = sorted(list(self.shards) + list(extra_shards), key=lambda s: s["first_item"])
        pos = 0
        for s in shards:
            if s["first_item"] != pos:
                raise ValueError(f"token table has a gap or overlap at item {pos}")
            pos += s["num_items"]
        meta = {"format": TABLE_FORMAT, "version": TABLE_VERSION, "num_items": pos, "tokens_per_item": self.q,
                "hidden": self.h, "dtype": "bfloat16", "shards": shards}
        with open(os.path.join(self.dir, "table.json"), "w") as f:
            json.dump(meta, f)
        if item_ids is not None:
            if len(item_ids) != pos:
                raise ValueError("len(item_ids) differs from the number of rows written")
            with open(os.path.join(self.dir, "item_ids.json"), "w") as f:
                json.dump(list(item_ids), f)


class TokenTable:
    """Memory-mapped view of a token table directory."""

    def __init__(self, directory: str):
        self.dir = directory
        with open(os.path.join(directory, "table.json")) as f:
            m = json.load(f)
        if m.get("format") != TABLE_FORMAT or m.get("version") != TABLE_VERSION or m.get("dtype") != "bfloat16":
            raise ValueError(f"{directory}: not a {TABLE_FORMAT} v{TABLE_VERSION} table")
        self.meta = m
        self.num_items, self.q, self.h = m["num_items"], m["tokens_per_item"], m["hidden"]
        self.shards = m["shards"]
        ids_path = os.path.join(directory, "item_ids.json")
        self.item_ids = json.load(open(ids_path)) if os.path.exists(ids_path) else None

    def shard_array(self, k: int) -> np.ndarray:
        s = self.shards[k]
        path = os.path.join(self.dir, s["file"])
        want = s["num_items"] * self.q * self.h * 2
        if os.path.getsize(path) != want:
            raise ValueError(f"{path}: size {os.path.getsize(path)} != {want}")
        return np.memmap(path, dtype=np.uint16, mode="r", shape=(s["num_items"], self.q, self.h))

    def read(self, lo: int = 0, hi: Optional[int] = None, device: Optional[torch.device] = None,
             chunk_items: int = 8192) -> torch.Tensor:
        """Rows [lo, hi) as a bf16 tensor; with a CUDA `device` the rows go through a pinned staging buffer in
        chunks of `chunk_items` (2 MiB-aligned large copies instead of one small copy per item)."""
        hi = self.num_items if hi is None else hi
        if not 0 <= lo <= hi <= self.num_items:
            raise IndexError((lo, hi, self.num_items))
        out = torch.empty(hi - lo, self.q, self.h, dtype=torch.bfloat16, device=device or "cpu")
        stage = None
        if out.is_cuda:
            stage = torch.empty(min(chunk_items, max(hi - lo, 1)), self.q, self.h, dtype=torch.bfloat16).pin_memory()
        for k, s in enumerate(self.shards):
            a, b = max(lo, s["first_item"]), min(hi, s["first_item"] + s["num_items"])
            if a >= b:
                continue
            arr = self.shard_array(k)
            for c0 in range(a, b, chunk_items):
                c1 = min(c0 + chunk_items, b)
                src = torch.from_numpy(np.array(arr[c0 - s["first_item"]:c1 - s["first_item"]]).view(np.int16))
                src = src.view(torch.bfloat16)
                if stage is None:
                    out[c0 - lo:c1 - lo].copy_(src)
                else:
                    torch.cuda.current_stream(out.device).synchronize()      # staging buffer free again
                    stage[:c1 - c0].copy_(src)
                    out[c0 - lo:c1 - lo].copy_(stage[:c1 - c0], non_blocking=True)
        if stage is not None:
            torch.cuda.current_stream(out.device).synchronize()
        return out


def pickle_to_token_table(pickle_path: str, directory: str) -> TokenTable:
    """Convert the reference's item-token pickle into the native table (ids kept in item_ids.json)."""
    ids, tok = load_item_tokens_pickle(pickle_path)
    w = TokenTableWriter(directory, tok.shape[1], tok.shape[2])
    w.write_shard(0, tok)
    w.close(item_ids=ids)
    return TokenTable(directory)


# --------------------------------------------------------------------------------------------- checkpoints
def save_item_checkpoint(path: str, model, field_names: Sequence[str]) -> None:
    """Write the reference's checkpoint dictionary (training/item_qformer_training.py:176-184):
    {'model_state_dict', 'config' (the BertConfig object), 'field_names'} - loadable by the reference's
    `load_qformer_model` (data_processing/qformer_inference.py:13-55) and `load_trained_model`
    (evaluation/evaluate_item_qformer.py:14-38) as well as by `load_item_checkpoint` below."""
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    torch.save({"model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                "config": model.config, "field_names": list(field_names)}, path)


def load_item_checkpoint(path: str, device=None):
    """The reference's loaders (qformer_inference.py:13-55 / evaluate_item_qformer.py:14-38) over the CUDA modules:
    rebuild the model from the stored config and field list, load the weights (same keys, strict), eval().
    Returns (model, field_names).  Raises ValueError when the checkpoint holds no 'field_names' (:21-22)."""
    from .modules import QFormerForItemRepresentation
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    cfg = ckpt["config"]
    field_names = ckpt.get("field_names")
    if field_names is None:
        raise ValueError("Checkpoint must contain 'field_names'")
    model = QFormerForItemRepresentation(
        hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
        num_attention_heads=cfg.num_attention_heads, intermediate_size=cfg.intermediate_size,
        num_query_tokens=cfg.query_length, field_embedding_dim=cfg.encoder_width, num_fields=len(field_names),
        dropout=cfg.hidden_dropout_prob)
    model.load_state_dict(ckpt["model_state_dict"])
    if device is not None:
        model = model.to(device)
    model.eval()
    return model, field_names
