"""ctypes binding of libunirec_b200.so (C ABI declared in include/unirec_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing, or a call
returns a non-zero code, a RuntimeError is raised (north_star: "no Triton, no multi-backend dispatch
and no CPU fallback").  Build the library with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C unirec_b200/csrc`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libunirec_b200.so")

_lib = None

# name -> (restype, argtypes); mirrors include/unirec_b200.h one to one
_SIGNATURES = {
    "unirec_abi_version": (c_int, []),
    "unirec_last_error": (c_char_p, []),
    "unirec_launch_count": (c_int64, []),
    "unirec_linear_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                   c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                   c_void_p]),
    "unirec_layernorm": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_float,
                                 c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p]),
    "unirec_attention": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                 c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_float,
                                 c_void_p]),
    "unirec_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "unirec_mean_tokens": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "unirec_field_projection": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64,
                                        c_int64, c_void_p]),
    "unirec_build_user_sequence": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "unirec_positional_encoding": (c_int, [c_void_p, c_int64, c_int64, c_void_p]),
    "unirec_inv_l2_norm": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int64, c_int64, c_float, c_void_p]),
    "unirec_score_topk_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "unirec_score_topk": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                  c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "unirec_topk_merge": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "unirec_gemm_general": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int,
                                    c_int64, c_int64, c_int64, c_int, c_void_p]),
    "unirec_gelu_forward": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "unirec_gelu_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "unirec_colsum": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "unirec_layernorm_backward": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_float,
                                          c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "unirec_layernorm_backward_fused": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_float,
                                                c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_uint32,
                                                c_uint64, c_uint32, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "unirec_attention_backward": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                          c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                          c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p]),
    "unirec_attention_dropout": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                         c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_float,
                                         c_uint32, c_uint64, c_uint32, c_void_p, c_void_p]),
    "unirec_attention_dropout_backward": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                                  c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                                  c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                                                  c_int64, c_float, c_uint32, c_uint64, c_uint32, c_void_p, c_void_p]),
    "unirec_dropout_add": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                   c_uint32, c_uint64, c_uint32, c_void_p, c_void_p]),
    "unirec_dropout_backward": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_uint32, c_uint64,
                                        c_uint32, c_void_p, c_void_p]),
    "unirec_list_scores": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                   c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "unirec_infonce_rank": (c_int, [c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "unirec_list_scores_backward": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p,
                                            c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p,
                                            c_float, c_void_p, c_void_p, c_void_p]),
    "unirec_inject_tokens": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int,
                                     c_int64, c_int64, c_void_p]),
    "unirec_inject_tokens_backward": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int64,
                                              c_void_p, c_int64, c_void_p]),
    "unirec_context_hidden": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                      c_void_p, c_int64, c_void_p, c_void_p]),
    "unirec_mwne_encode": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_void_p, c_int, c_void_p]),
    "unirec_linear_ln_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                      c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int, c_float, c_int64, c_void_p]),
    "unirec_linear_ln_stats_parts": (c_int64, [c_int64]),
    "unirec_kv_attention_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "unirec_kv_attention_fused": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                          c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                          c_int64, c_float, c_void_p]),
    "unirec_linear_gather_bf16": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                          c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                          c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "unirec_reconstruction_metrics": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p,
                                              c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load():
    """Load the shared library (once) and attach argument types.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built and unirec_b200 has no CPU fallback. "
            "Run `make -C unirec_b200/csrc` (or __graft_entry__.build()).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().unirec_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().unirec_launch_count())
