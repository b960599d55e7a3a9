"""CPU: the drop-in boundary - C-ABI symbols, header/binding agreement, module contract, loud failure
without a GPU.  No compute kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "unirec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(unirec_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from unirec_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the extension first: make -C unirec_b200/csrc"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert declared, "no symbols parsed from include/unirec_b200.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/unirec_b200.h but not exported"
    # the ctypes binding covers exactly the declared surface
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert _lib.load().unirec_abi_version() == 3


def test_workspace_query_is_host_only():
    from unirec_b200 import _lib
    lib = _lib.load()
    assert lib.unirec_score_topk_workspace_bytes(4096, 1_000_000, 100) > 0
    assert lib.unirec_score_topk_workspace_bytes(4096, 1_000_000, 1000) < 0   # k > 128 unsupported
    assert b"k" in lib.unirec_last_error()


def test_state_dict_keys_match_reference_layout():
    from unirec_b200 import synth
    from unirec_b200.modules import QFormerForItemRepresentation, UserQFormer
    kw = dict(hidden_size=256, num_hidden_layers=4, num_attention_heads=4, intermediate_size=512)
    item = QFormerForItemRepresentation(field_embedding_dim=256, num_fields=6, **kw)
    ref_keys = set(synth.item_qformer_state_dict(hidden=256, layers=4, inter=512, field_dim=256, num_fields=6))
    assert set(item.state_dict()) == ref_keys
    # cross-attention only in even layers for the item model (cross_attention_freq=2), all layers for the user model
    assert "qformer.encoder.layer.0.crossattention.self.key.weight" in ref_keys
    assert "qformer.encoder.layer.1.crossattention.self.key.weight" not in ref_keys
    user = UserQFormer(input_embedding_dim=256, num_item_tokens_to_predict=8, **kw)
    ukeys = set(synth.user_qformer_state_dict(hidden=256, layers=4, inter=512, input_dim=256, num_predict=8))
    assert set(user.state_dict()) == ukeys
    assert "qformer.encoder.layer.1.crossattention.self.key.weight" in ukeys
    # dead tensors of the reference checkpoint round-trip
    for k in ("qformer.embeddings.word_embeddings.weight", "qformer.embeddings.position_ids",
              "qformer.encoder.layer.0.intermediate.dense.weight", "qformer.encoder.layer.0.output.LayerNorm.bias"):
        assert k in ref_keys
    sd = synth.item_qformer_state_dict(hidden=256, layers=4, inter=512, field_dim=256, num_fields=6)
    item.load_state_dict(sd, strict=True)
    for k, v in item.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_constructor_contract_and_attributes():
    from unirec_b200.modules import QFormerForItemRepresentation, UserQFormer
    with pytest.raises(ValueError):
        QFormerForItemRepresentation()                      # num_fields is required (qformer_utils.py:21)
    with pytest.raises(ValueError):
        QFormerForItemRepresentation(hidden_size=250, num_attention_heads=4, num_fields=3)
    m = QFormerForItemRepresentation(hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                                     intermediate_size=512, field_embedding_dim=256, num_fields=6, dropout=0.2)
    c = m.config
    assert (c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.intermediate_size) == (256, 2, 4, 512)
    assert (c.query_length, c.encoder_width, c.hidden_dropout_prob, c.cross_attention_freq) == (32, 256, 0.2, 2)
    assert m.num_query_tokens == 32 and tuple(m.query_embeddings.shape) == (1, 32, 256)
    u = UserQFormer(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                    input_embedding_dim=256, num_item_tokens_to_predict=8)
    assert u.num_query_tokens == 64 and u.config.cross_attention_freq == 1
    assert u.prediction_head[3].out_features == 8 * 256 and u.prediction_head[2].eps == 1e-5
    # reference init rule: zero Linear bias, N(0, 0.02) weights
    w = m.qformer.encoder.layer[0].attention.self.query.weight
    assert 0.015 < float(w.detach().std()) < 0.025
    assert float(m.qformer.encoder.layer[0].attention.self.query.bias.detach().abs().max()) == 0


def test_no_cpu_fallback():
    from unirec_b200 import ops
    from unirec_b200.modules import QFormerForItemRepresentation
    m = QFormerForItemRepresentation(hidden_size=256, num_hidden_layers=2, num_attention_heads=4,
                                     intermediate_size=512, field_embedding_dim=256, num_fields=6).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(2, 6, 256))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(torch.randn(4, 64).bfloat16(), torch.randn(8, 64).bfloat16())
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.score_topk(torch.randn(4, 64).bfloat16(), torch.randn(8, 64).bfloat16(), 2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unirec_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("qformer_oracle", "oracle") or "import oracle" not in src
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_shard_range_partitions_exactly():
    from unirec_b200.pipeline import shard_range
    for n in (0, 1, 7, 1000, 1_000_000, 1_000_003):
        for g in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, g) for r in range(g)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_torch_custom_ops_are_registered_with_fake_impls_and_no_cpu_kernel():
    """The torch custom-op layer (unirec_b200/torch_ops.py): every op of the namespace infers shapes / dtypes on meta
    tensors (what FakeTensorMode and torch.export use) and refuses CPU tensors - there is no CPU kernel."""
    import pytest
    import torch
    import unirec_b200.torch_ops as T
    ns = torch.ops.unirec_b200
    for name in T.OPS:
        assert hasattr(ns, name), name

    def m(*shape, dt=torch.bfloat16):
        return torch.empty(*shape, device="meta", dtype=dt)

    f32, i64 = torch.float32, torch.int64
    y = ns.linear(m(64, 1024), m(4096, 1024), m(4096, dt=f32), None, 1, False)
    assert tuple(y.shape) == (64, 4096) and y.dtype == torch.bfloat16
    y = ns.linear(m(2, 32, 1024), m(1024, 1024), None, m(2, 32, 1024), 2, True)
    assert tuple(y.shape) == (2, 32, 1024) and y.dtype == f32
    y, st = ns.linear_ln(m(64, 4096), m(1024, 4096), m(1024, dt=f32), m(64, 1024), 2, None, None, m(64, 8, 2, dt=f32),
                         m(1024, dt=f32), m(1024, dt=f32), True, 1e-12, 1024)
    assert tuple(y.shape) == (64, 1024) and y.dtype == torch.bfloat16 and tuple(st.shape) == (64, 8, 2) and st.dtype == f32
    y, st = ns.linear_ln(m(64, 1024), m(4096, 1024), m(4096, dt=f32), None, 1, m(64, 8, 2, dt=f32), m(4096, dt=f32), None,
                         None, None, False, 1e-12, 1024)
    assert tuple(y.shape) == (64, 4096) and st.numel() == 0
    assert tuple(ns.layernorm(m(32, 1024, dt=f32), m(1024, dt=f32), m(1024, dt=f32), 1e-12, None, 256, 32, False).shape) == (256, 1024)
    assert tuple(ns.attention(m(64, 1024), m(28, 1024), m(28, 1024), m(2, 14, dt=f32), 2, 16, 32, 14, False).shape) == (64, 1024)
    assert ns.cast_bf16(m(3, 5, dt=f32)).dtype == torch.bfloat16
    assert tuple(ns.mean_tokens(m(7, 32, 1024), True).shape) == (7, 1024)
    assert tuple(ns.field_projection(m(7, 32, 1024), m(14, 32, dt=f32), m(14, dt=f32), True).shape) == (7, 14, 1024)
    seq, mask = ns.build_user_sequence(m(100, 32, 1024), m(4, 50, dt=i64), m(4, dt=torch.int32), None)
    assert tuple(seq.shape) == (4, 1600, 1024) and tuple(mask.shape) == (4, 1600) and mask.dtype == f32
    assert tuple(ns.inv_l2_norm(m(9, 1024), 1e-12).shape) == (9,)
    s, i = ns.score_topk(m(8, 1024), m(1000, 1024), 100, None, None, 0)
    assert tuple(s.shape) == (8, 100) and s.dtype == f32 and i.dtype == i64
    s, i = ns.topk_merge(m(4, 8, 100, dt=f32), m(4, 8, 100, dt=i64))
    assert tuple(s.shape) == (8, 100) and tuple(i.shape) == (8, 100)
    sims, inv = ns.list_scores(m(8, 256, dt=f32), m(8, 256, dt=f32), m(8, 20, 256, dt=f32), None, None, 0, 1e-12)
    assert tuple(sims.shape) == (8, 21) and tuple(inv.shape) == (8, 21)
    sims, _ = ns.list_scores(m(8, 256, dt=f32), m(8, 256, dt=f32), m(90, 256, dt=f32), None, m(9, dt=i64), 33, 1e-12)
    assert tuple(sims.shape) == (8, 34)
    loss, rank = ns.infonce_rank(m(8, 21, dt=f32), 0.07)
    assert tuple(loss.shape) == (8,) and rank.dtype == torch.int32
    with pytest.raises(NotImplementedError):
        ns.linear(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(256, 64, dtype=torch.bfloat16), None, None, 0, False)


def test_new_entry_points_refuse_cpu_inputs():
    """The joint-trainer hooks, the evaluation loop and the streamed generation loop have no CPU path either: a CPU
    tensor / CPU model raises instead of computing something somewhere else."""
    import pytest
    import torch
    from unirec_b200.evaluation import evaluate_reconstruction
    from unirec_b200.joint import InfoNCELoss, inject_history_tokens
    from unirec_b200.modules import QFormerForItemRepresentation, UserQFormer
    from unirec_b200.pipeline import generate_item_tokens_streamed
    u, n = torch.randn(4, 64), torch.randn(4, 3, 64)
    with pytest.raises(RuntimeError):
        InfoNCELoss()(u, u, n)
    with pytest.raises(RuntimeError):
        inject_history_tokens(torch.zeros(1, 4, 64), torch.zeros(1, 4, dtype=torch.long), torch.arange(2).view(1, 2),
                              torch.zeros(1, 1, 2, 64))
    model = QFormerForItemRepresentation(hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                                         intermediate_size=128, num_query_tokens=8, field_embedding_dim=128, num_fields=3)
    with pytest.raises(RuntimeError):
        evaluate_reconstruction(model.eval(), torch.zeros(2, 3, 128), torch.ones(2, 3))
    with pytest.raises(RuntimeError):
        generate_item_tokens_streamed(model.eval(), torch.zeros(2, 3, 128), None, torch.zeros(2, 8, 128, dtype=torch.bfloat16))
    um = UserQFormer(hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=128,
                     num_query_tokens=8, input_embedding_dim=128, num_item_tokens_to_predict=2).eval()
    with pytest.raises(RuntimeError):
        um.encode_queries_from_history(torch.zeros(5, 16, 128, dtype=torch.bfloat16), torch.zeros(1, 2, dtype=torch.long),
                                       torch.ones(1, dtype=torch.int32))


def test_plain_c_program_binds_the_abi(tmp_path):
    """include/unirec_b200.h compiles as C11 and a C program (examples/c_abi_probe.c) loads the library without Python,
    finds the entry points, runs the host-only ones and gets an error CODE (not a crash) from a bad call."""
    import shutil
    import subprocess
    from unirec_b200 import _lib
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    exe = str(tmp_path / "c_abi_probe")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_probe.c"), "-ldl", "-o", exe], check=True)
    out = subprocess.run([exe, _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout
    assert out.startswith("abi 3,") and "rc 1" in out
