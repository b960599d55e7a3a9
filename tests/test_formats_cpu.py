"""On-disk formats: files written here must load with the reference's own reading code
(models/qformer_utils.py:121-137; data_processing/qformer_inference.py:210-211) and files written the way the
reference writes them (:139-145; qformer_inference.py:163-173) must load here."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from unirec_b200 import formats


def test_field_cache_round_trip_and_reference_layout(tmp_path):
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(7, 5, 16, generator=g)
    mask = (torch.rand(7, 5, generator=g) < 0.8).long()
    names = sorted(["title", "price", "brand", "main_image", "category"])
    d = str(tmp_path / "cache")
    formats.save_field_cache(d, emb, mask, names)
    # the reference's loader: torch.load of two {idx: tensor} dicts + fields.json equality check
    e = torch.load(os.path.join(d, "embeddings.pt"), weights_only=False)
    m = torch.load(os.path.join(d, "masks.pt"), weights_only=False)
    assert json.load(open(os.path.join(d, "fields.json"))) == names
    assert set(e) == set(range(7)) and e[3].dtype == torch.float32 and m[3].dtype == torch.long
    assert torch.equal(e[3], emb[3]) and torch.equal(m[6], mask[6])
    assert os.path.getsize(os.path.join(d, "embeddings.pt")) < 4 * emb.numel() * 4    # entries are not views
    got = formats.load_field_cache(d, expected_fields=names)
    assert got is not None and torch.equal(got[0], emb) and torch.equal(got[1], mask) and got[2] == names
    assert formats.load_field_cache(d, expected_fields=names[:-1] + ["other"]) is None     # outdated cache
    assert formats.load_field_cache(str(tmp_path / "missing")) is None


def test_field_cache_written_by_reference_code_loads(tmp_path):
    """Write the files exactly like QFormerDataset._precompute/_save_cache (dicts filled sample by sample)."""
    d = tmp_path / "ref_cache"
    d.mkdir()
    rng = np.random.default_rng(1)
    cache, masks = {}, {}
    for idx in (2, 0, 1):          # insertion order is not index order
        arr = [rng.standard_normal(8).astype(np.float32) for _ in range(3)]
        arr[1][:] = 0.0 if idx == 1 else arr[1]
        cache[idx] = torch.tensor(np.array(arr), dtype=torch.float32)
        masks[idx] = torch.tensor([1 if np.any(a) else 0 for a in arr], dtype=torch.long)
    torch.save(cache, d / "embeddings.pt")
    torch.save(masks, d / "masks.pt")
    json.dump(["a", "b", "c"], open(d / "fields.json", "w"))
    emb, msk, names = formats.load_field_cache(str(d), expected_fields=["a", "b", "c"])
    assert emb.shape == (3, 3, 8) and torch.equal(emb[1], cache[1]) and msk[1].tolist() == [1, 0, 1]


def test_item_token_pickle_round_trip(tmp_path):
    g = torch.Generator().manual_seed(2)
    tok = torch.randn(9, 4, 32, generator=g)
    ids = [f"B00{i:05d}" for i in range(9)]
    p = str(tmp_path / "out" / "tokens.pkl")
    formats.save_item_tokens_pickle(p, ids, tok)
    saved = pickle.load(open(p, "rb"))                                       # the reference's reader
    assert list(saved) == ids and saved[ids[4]].dtype == np.float32 and saved[ids[4]].shape == (4, 32)
    np.testing.assert_array_equal(saved[ids[4]], tok[4].numpy())
    ids2, t2 = formats.load_item_tokens_pickle(p, dtype=torch.float32)
    assert ids2 == ids and torch.equal(t2, tok)


def test_token_table_shards_memmap_and_gap_detection(tmp_path):
    g = torch.Generator().manual_seed(3)
    tok = torch.randn(50, 4, 16, generator=g).to(torch.bfloat16)
    d = str(tmp_path / "table")
    w = formats.TokenTableWriter(d, 4, 16)
    w.write_shard(20, tok[20:50])            # shards may be written out of order (one per rank)
    w.write_shard(0, tok[:20])
    w.close(item_ids=list(range(100, 150)))
    t = formats.TokenTable(d)
    assert (t.num_items, t.q, t.h) == (50, 4, 16) and t.item_ids[7] == 107
    assert torch.equal(t.read(), tok)
    assert torch.equal(t.read(15, 33, chunk_items=4), tok[15:33])            # crosses the shard boundary
    assert t.read(5, 5).shape == (0, 4, 16)
    with pytest.raises(IndexError):
        t.read(0, 51)
    w2 = formats.TokenTableWriter(str(tmp_path / "bad"), 4, 16)
    w2.write_shard(0, tok[:10])
    w2.write_shard(12, tok[12:20])
    with pytest.raises(ValueError):
        w2.close()
    # truncated shard file is detected
    path = os.path.join(d, t.shards[0]["file"])
    with open(path, "r+b") as f:
        f.truncate(os.path.getsize(path) - 2)
    with pytest.raises(ValueError):
        formats.TokenTable(d).read()


def test_pickle_to_token_table(tmp_path):
    tok = torch.randn(6, 2, 8, generator=torch.Generator().manual_seed(4))
    p = str(tmp_path / "tokens.pkl")
    formats.save_item_tokens_pickle(p, ["x", "y", "z", "u", "v", "w"], tok)
    t = formats.pickle_to_token_table(p, str(tmp_path / "tbl"))
    assert t.item_ids == ["x", "y", "z", "u", "v", "w"]
    assert torch.equal(t.read(), tok.to(torch.bfloat16))


def test_item_checkpoint_round_trip(tmp_path):
    """The reference's checkpoint dictionary ({'model_state_dict', 'config', 'field_names'},
    training/item_qformer_training.py:176-184) written and read back: same config attributes the reference's loaders
    read (qformer_inference.py:36-45), same weights, strict key match; missing 'field_names' raises ValueError."""
    import pytest
    import torch
    from unirec_b200 import formats
    from unirec_b200.modules import QFormerForItemRepresentation
    fields = ["brand", "categories", "main_image", "price", "title"]
    model = QFormerForItemRepresentation(hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                         intermediate_size=256, num_query_tokens=8, field_embedding_dim=96,
                                         num_fields=len(fields), dropout=0.15)
    path = str(tmp_path / "ckpt" / "best_qformer_model.pth")
    formats.save_item_checkpoint(path, model, fields)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert set(raw) == {"model_state_dict", "config", "field_names"}
    cfg = raw["config"]
    assert (cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size, cfg.query_length,
            cfg.encoder_width, cfg.hidden_dropout_prob) == (128, 2, 2, 256, 8, 96, 0.15)
    loaded, names = formats.load_item_checkpoint(path)
    assert names == fields and not loaded.training and loaded.num_query_tokens == 8
    sd0, sd1 = model.state_dict(), loaded.state_dict()
    assert list(sd0) == list(sd1)
    for k in sd0:
        assert torch.equal(sd0[k], sd1[k]), k
    del raw["field_names"]
    torch.save(raw, path)
    with pytest.raises(ValueError):
        formats.load_item_checkpoint(path)
