"""Tensor-level wrappers over the C ABI (include/unirec_b200.h).

PyTorch is used for what it is good at here - owning device memory and the current CUDA stream; every
arithmetic op below is one launch of a hand-written sm_100a kernel in libunirec_b200.so.  Inputs
must be CUDA tensors; anything else raises (no CPU fallback).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL = 0, 1, 2

# Optional per-launch device timing (bench.py's roofline): when a list is installed, every wrapped
# launch is bracketed by CUDA events on the launching stream and (kind, work, start, end) is appended;
# `work` is the launch's ALGORITHMIC flops (GEMM, scoring) or bytes (attention).
_timing = None


def start_timing():
    global _timing
    _timing = []


def stop_timing():
    """Returns {kind: (launches, total_work, total_ms)} plus, per tagged launch shape, {"kind:tag": (...)}; call after
    torch.cuda.synchronize()."""
    global _timing
    rec, _timing = _timing or [], None
    out = {}
    for kind, tag, work, e0, e1 in rec:
        dt = e0.elapsed_time(e1)
        for key in ((kind,) if tag is None else (kind, f"{kind}:{tag}")):
            n, w, ms = out.get(key, (0, 0.0, 0.0))
            out[key] = (n + 1, w + work, ms + dt)
    return out


class _Timed:
    def __init__(self, kind, work, tag=None):
        self.kind, self.work, self.tag = kind, work, tag

    def __enter__(self):
        if _timing is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.stream = torch.cuda.current_stream(torch._C._cuda_getDevice())
            self.e0.record(self.stream)
        return self

    def __exit__(self, *exc):
        if _timing is not None:
            self.e1.record(self.stream)
            _timing.append((self.kind, self.tag, float(self.work), self.e0, self.e1))
        return False


def _stream() -> int:
    # raw cudaStream_t of torch's current stream.  torch.cuda.current_stream() with no device argument goes through
    # torch.cuda.is_available() -> cudaGetDeviceCount on every call (cProfile: the largest host cost of a training
    # step); the two C calls below do not.
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _req(t: torch.Tensor, dtype, name: str):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (unirec_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.dim() >= 1 and t.stride(-1) != 1 and t.shape[-1] != 1:
        raise RuntimeError(f"{name}: last dimension must be contiguous")


def _rows2d(t: torch.Tensor, name: str) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a tensor viewed as 2-D [rows, cols] with a uniform row stride."""
    if t.dim() == 2:
        return t.shape[0], t.shape[1], t.stride(0)
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: tensors with more than 2 dims must be contiguous")
    return t.numel() // t.shape[-1], t.shape[-1], t.shape[-1]


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def linear(a: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
           epilogue: int = EPI_BIAS, residual: Optional[torch.Tensor] = None, res_row_mod: int = 0,
           out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
           block_n: int = 0, max_ctas: int = 0) -> torch.Tensor:
    """out = epilogue(a @ weight.T + bias); a [.., K] bf16, weight [N, K] bf16, bias [N] fp32."""
    _req(a, torch.bfloat16, "linear.a")
    _req(weight, torch.bfloat16, "linear.weight")
    M, K, lda = _rows2d(a, "linear.a")
    N, K2 = weight.shape
    if K2 != K:
        raise RuntimeError(f"linear: inner dims differ ({K} vs {K2})")
    if bias is not None:
        _req(bias, torch.float32, "linear.bias")
    if out is None:
        out = torch.empty(*a.shape[:-1], N, device=a.device, dtype=out_dtype)
    else:
        _req(out, out.dtype, "linear.out")
    _, No, ldo = _rows2d(out, "linear.out")
    if No != N:
        raise RuntimeError("linear: out has the wrong width")
    ldr = 0
    if residual is not None:
        _req(residual, torch.bfloat16, "linear.residual")
        _, _, ldr = _rows2d(residual, "linear.residual")
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise RuntimeError("linear: out must be bf16 or fp32")
    with _Timed("gemm", 2.0 * M * N * K, f"{M}x{N}x{K}"):
        rc = _lib.load().unirec_linear_bf16(a.data_ptr(), lda, weight.data_ptr(), weight.stride(0), _ptr(bias),
                                            _ptr(residual), ldr, res_row_mod, out.data_ptr(), ldo,
                                            1 if out.dtype == torch.float32 else 0, M, N, K, epilogue, block_n,
                                            max_ctas, _stream())
    _lib.check(rc, "unirec_linear_bf16")
    return out


def linear_ln(a: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, *, epilogue: int = EPI_BIAS,
              residual: Optional[torch.Tensor] = None, ln_in: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
              ln_res: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
              stats_out: Optional[torch.Tensor] = None, eps: float = 1e-12, hidden: Optional[int] = None) -> torch.Tensor:
    """nn.Linear with the LayerNorms around it folded in (C ABI unirec_linear_ln_bf16; bf16 in / out, N % 256 == 0):
      ln_in  = (stats fp32 [M, P, 2], c fp32 [N]): `a` is the INPUT of a LayerNorm (never materialised); `weight` / `bias`
               are the gamma-scaled weight W' = W o gamma and b' = b + W beta, c = row sums of W' (fold_layernorm_weights);
      ln_res = (stats fp32 [M, P, 2], gamma fp32 [N], beta fp32 [N]): `residual` is the input of a LayerNorm and enters
               normalised;
      stats_out fp32 [M, parts(N), 2] (see `ln_stats_buffer`): receives, per row, the (sum, sum of squares) of every
               column piece of the bf16 output - the `stats` of the calls that consume this output."""
    _req(a, torch.bfloat16, "linear_ln.a")
    _req(weight, torch.bfloat16, "linear_ln.weight")
    _req(bias, torch.float32, "linear_ln.bias")
    M, K, lda = _rows2d(a, "linear_ln.a")
    N, K2 = weight.shape
    if K2 != K:
        raise RuntimeError(f"linear_ln: inner dims differ ({K} vs {K2})")
    out = torch.empty(*a.shape[:-1], N, device=a.device, dtype=torch.bfloat16)
    ldr = 0
    if residual is not None:
        _req(residual, torch.bfloat16, "linear_ln.residual")
        _, _, ldr = _rows2d(residual, "linear_ln.residual")
    parts = 0
    for t, n in ((ln_in, "ln_in"), (ln_res, "ln_res")):
        if t is not None:
            for x in t:
                _req(x, torch.float32, f"linear_ln.{n}")
            st = t[0]
            if st.dim() != 3 or st.shape[0] != M or st.shape[2] != 2 or not st.is_contiguous():
                raise RuntimeError(f"linear_ln.{n}: statistics must be contiguous fp32 [M, parts, 2]")
            if parts and st.shape[1] != parts:
                raise RuntimeError("linear_ln: ln_in and ln_res statistics must have the same number of partials")
            parts = int(st.shape[1])
    if stats_out is not None:
        _req(stats_out, torch.float32, "linear_ln.stats_out")
        if tuple(stats_out.shape) != (M, ln_stats_parts(N), 2) or not stats_out.is_contiguous():
            raise RuntimeError("linear_ln.stats_out must be contiguous fp32 [M, ln_stats_parts(N), 2]")
    hid = int(hidden) if hidden is not None else (K if ln_in is not None else N)
    with _Timed("gemm", 2.0 * M * N * K, f"{M}x{N}x{K}"):
        rc = _lib.load().unirec_linear_ln_bf16(
            a.data_ptr(), lda, weight.data_ptr(), weight.stride(0), bias.data_ptr(), _ptr(residual), ldr, out.data_ptr(),
            out.stride(-2) if out.dim() >= 2 else N, M, N, K, epilogue,
            _ptr(ln_in[0]) if ln_in is not None else None, _ptr(ln_in[1]) if ln_in is not None else None,
            _ptr(ln_res[0]) if ln_res is not None else None, _ptr(ln_res[1]) if ln_res is not None else None,
            _ptr(ln_res[2]) if ln_res is not None else None, _ptr(stats_out), parts, float(eps), hid, _stream())
    _lib.check(rc, "unirec_linear_ln_bf16")
    return out


def ln_stats_parts(width: int) -> int:
    """Partials per row that `linear_ln(..., stats_out=)` writes for `width` output columns (one per epilogue warp part)."""
    return int(_lib.load().unirec_linear_ln_stats_parts(int(width)))


def ln_stats_buffer(rows: int, width: int, device, count: Optional[int] = None) -> torch.Tensor:
    """Uninitialised statistics buffer(s) for `linear_ln(..., stats_out=)` of a GEMM with `width` output columns:
    fp32 [rows, parts, 2], or [count, rows, parts, 2], parts = ln_stats_parts(width)."""
    shape = (rows, ln_stats_parts(width), 2)
    return torch.empty(shape if count is None else (count,) + shape, device=device, dtype=torch.float32)


def fold_layernorm_weights(weight: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor):
    """(W' bf16 [N, K], b' fp32 [N], c fp32 [N]) of a Linear that consumes LayerNorm(x; gamma, beta):
    Linear(LN(x)) = rstd (x W'^T) - mu rstd c + b' with W' = W o gamma (rounded to bf16, what the tensor cores multiply),
    c = row sums of that rounded W' (so that the mean term cancels exactly what was accumulated), b' = b + W beta."""
    w = weight.detach().float()
    wp = (w * gamma.detach().float()[None, :]).to(torch.bfloat16).contiguous()
    b = w @ beta.detach().float()
    if bias is not None:
        b = b + bias.detach().float()
    return wp, b.contiguous(), wp.float().sum(dim=1).contiguous()


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *,
              residual: Optional[torch.Tensor] = None, rows: Optional[int] = None, in_row_mod: int = 0,
              out_dtype: torch.dtype = torch.bfloat16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm(x [+ residual]) over the last dim.  With in_row_mod > 0 the input has in_row_mod rows
    that are broadcast over `rows` output rows."""
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise RuntimeError("layernorm: x must be bf16 or fp32")
    _req(x, x.dtype, "layernorm.x")
    _req(gamma, torch.float32, "layernorm.gamma")
    _req(beta, torch.float32, "layernorm.beta")
    xr, H, ldx = _rows2d(x, "layernorm.x")
    if rows is None:
        rows = xr
    if out is None:
        shape = (rows, H) if (in_row_mod > 0 or x.dim() == 2) else x.shape
        out = torch.empty(shape, device=x.device, dtype=out_dtype)
    _, _, ldo = _rows2d(out, "layernorm.out")
    ldres = 0
    if residual is not None:
        _req(residual, torch.bfloat16, "layernorm.residual")
        _, _, ldres = _rows2d(residual, "layernorm.residual")
    rc = _lib.load().unirec_layernorm(x.data_ptr(), 1 if x.dtype == torch.float32 else 0, ldx, in_row_mod,
                                      _ptr(residual), ldres, gamma.data_ptr(), beta.data_ptr(), float(eps),
                                      out.data_ptr(), 1 if out.dtype == torch.float32 else 0, ldo, rows, H, _stream())
    _lib.check(rc, "unirec_layernorm")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, batch: int, num_heads: int, nq: int, nk: int,
              key_mask: Optional[torch.Tensor] = None, q_broadcast: bool = False,
              out: Optional[torch.Tensor] = None, dropout: Optional[Tuple[int, int, int]] = None) -> torch.Tensor:
    """Multi-head attention with head_dim 64.  q/k/v are 2-D row views (possibly column slices of a fused
    projection buffer): q [batch*nq (or nq if q_broadcast), heads*64], k/v [batch*nk, heads*64].
    dropout = (thr16, seed, site[, seed_offset]): train-mode dropout of the probabilities (see dropout_threshold,
    _drop_args)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, torch.bfloat16, f"attention.{n}")
        if t.dim() != 2:
            raise RuntimeError("attention: q/k/v must be 2-D row views")
    hd = num_heads * 64
    if q.shape[1] != hd or k.shape[1] != hd or v.shape[1] != hd:
        raise RuntimeError("attention: width must be num_heads*64")
    if out is None:
        out = torch.empty(batch * nq, hd, device=q.device, dtype=torch.bfloat16)
    if key_mask is not None:
        _req(key_mask, torch.float32, "attention.key_mask")
        if tuple(key_mask.shape) != (batch, nk) or not key_mask.is_contiguous():
            raise RuntimeError("attention: key_mask must be contiguous [batch, nk]")
    # algorithmic bytes: Q, K, V read once + context written once (bf16)
    with _Timed("attention", 2.0 * hd * batch * ((1 if q_broadcast else 1) * nq + 2 * nk + nq), f"{batch}x{nq}x{nk}"):
        if dropout is not None and dropout[0] > 0:
            thr16, seed, site, off = _drop_args(dropout)
            rc = _lib.load().unirec_attention_dropout(q.data_ptr(), q.stride(0), 0 if q_broadcast else nq, k.data_ptr(),
                                                      k.stride(0), v.data_ptr(), v.stride(0), nk, _ptr(key_mask),
                                                      out.data_ptr(), out.stride(0), batch, num_heads, nq, nk, 64, 0.125,
                                                      thr16, seed, site, off, _stream())
        else:
            rc = _lib.load().unirec_attention(q.data_ptr(), q.stride(0), 0 if q_broadcast else nq, k.data_ptr(),
                                              k.stride(0), v.data_ptr(), v.stride(0), nk, _ptr(key_mask), out.data_ptr(),
                                              out.stride(0), batch, num_heads, nq, nk, 64, 0.125, _stream())
    _lib.check(rc, "unirec_attention")
    return out


def pack_kv_weights(wk: torch.Tensor, wv: torch.Tensor) -> torch.Tensor:
    """[H, E] key and value weights -> the [2 H, E] bf16 operand of `kv_attention`: per head PAIR j, the 128 key rows of
    heads 2j, 2j+1 followed by their 128 value rows (one 256-column tile of the fused kernel = one head pair's K | V)."""
    H, E = wk.shape
    if H % 128 != 0 or tuple(wv.shape) != (H, E):
        raise RuntimeError("pack_kv_weights: needs [H, E] key / value weights with an even number of 64-wide heads")
    k = wk.detach().reshape(H // 128, 128, E)
    v = wv.detach().reshape(H // 128, 128, E)
    return torch.cat([k, v], dim=1).reshape(2 * H, E).to(torch.bfloat16).contiguous()


def kv_attention_supported(nq: int, nk: int, num_heads: int, enc_width: int) -> bool:
    return nq == 64 and nk >= 64 and nk % 64 == 0 and num_heads % 2 == 0 and enc_width % 64 == 0


def kv_attention(x: torch.Tensor, w_packed: torch.Tensor, q: torch.Tensor, v_bias: Optional[torch.Tensor], *, batch: int,
                 num_heads: int, nk: int, key_mask: Optional[torch.Tensor] = None, q_broadcast: bool = False,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Cross-attention of 64 queries per user over nk keys with the key / value projection fused in (no K / V in memory):
    x bf16 [batch * nk, E] encoder states, w_packed = pack_kv_weights(Wk, Wv), q bf16 [batch * 64 (or 64), H] projected
    queries, v_bias fp32 [H].  Returns bf16 [batch * 64, H]."""
    _req(x, torch.bfloat16, "kv_attention.x")
    _req(w_packed, torch.bfloat16, "kv_attention.w_packed")
    _req(q, torch.bfloat16, "kv_attention.q")
    if x.dim() != 2 or q.dim() != 2 or w_packed.dim() != 2:
        raise RuntimeError("kv_attention: x, q and w_packed must be 2-D row views")
    H = num_heads * 64
    E = x.shape[1]
    if x.shape[0] != batch * nk or tuple(w_packed.shape) != (2 * H, E) or q.shape[1] != H or \
            q.shape[0] != (64 if q_broadcast else batch * 64):
        raise RuntimeError("kv_attention: shapes do not match (x [batch * nk, E], w_packed [2 H, E], q [batch * 64, H])")
    if not kv_attention_supported(64, nk, num_heads, E):
        raise RuntimeError(f"kv_attention: needs nk % 64 == 0, an even head count, E % 64 == 0 (nk={nk} heads={num_heads})")
    if v_bias is not None:
        _req(v_bias, torch.float32, "kv_attention.v_bias")
    if key_mask is not None:
        _req(key_mask, torch.float32, "kv_attention.key_mask")
        if tuple(key_mask.shape) != (batch, nk) or not key_mask.is_contiguous():
            raise RuntimeError("kv_attention: key_mask must be contiguous [batch, nk]")
    if out is None:
        out = torch.empty(batch * 64, H, device=x.device, dtype=torch.bfloat16)
    lib = _lib.load()
    ws_bytes = int(lib.unirec_kv_attention_workspace_bytes(batch, num_heads))
    ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
    flops = 2.0 * batch * nk * (2 * H) * E + 4.0 * batch * num_heads * 64 * nk * 64
    with _Timed("kv_attention", flops, f"{batch * nk}x{2 * H}x{E}"):
        rc = lib.unirec_kv_attention_fused(x.data_ptr(), x.stride(0), w_packed.data_ptr(), w_packed.stride(0), q.data_ptr(),
                                           q.stride(0), 0 if q_broadcast else 64, _ptr(key_mask), _ptr(v_bias),
                                           out.data_ptr(), out.stride(0), ws.data_ptr(), ws_bytes, batch, nk, num_heads, E,
                                           0.125, _stream())
    _lib.check(rc, "unirec_kv_attention_fused")
    return out


def _drop_args(dropout):
    """(thr16, seed, site[, seed_offset]) -> (thr16, seed, site, device pointer of the uint64 seed offset or None).
    seed_offset: int64 CUDA tensor with one element, added to `seed` by the kernels when they run (CUDA-graph replays)."""
    if dropout is None:
        return 0, 0, 0, None
    off = dropout[3] if len(dropout) > 3 else None
    if off is not None:
        _req(off, torch.int64, "dropout.seed_offset")
        if off.numel() != 1:
            raise RuntimeError("dropout.seed_offset must hold one int64")
    return dropout[0], dropout[1], dropout[2], _ptr(off)


def dropout_threshold(p: float) -> int:
    """thr16 = round(p * 65536): an element is dropped iff its 16-bit Philox value is below thr16."""
    if not 0.0 <= p < 1.0:
        raise ValueError(f"dropout probability has to be in [0, 1), got {p}")
    return int(round(p * 65536.0))


def dropout_add(x: torch.Tensor, residual: Optional[torch.Tensor], dropout: Tuple[int, int, int], *,
                rows: Optional[int] = None, x_row_mod: int = 0) -> torch.Tensor:
    """out = dropout(x) [+ residual]; bf16 [rows, H].  With x_row_mod > 0, x has x_row_mod rows broadcast over `rows`."""
    _req(x, torch.bfloat16, "dropout_add.x")
    xr, H, ldx = _rows2d(x, "dropout_add.x")
    rows = xr if rows is None else rows
    ldres = 0
    if residual is not None:
        _req(residual, torch.bfloat16, "dropout_add.residual")
        _, _, ldres = _rows2d(residual, "dropout_add.residual")
    out = torch.empty(rows, H, device=x.device, dtype=torch.bfloat16)
    thr16, seed, site, off = _drop_args(dropout)
    rc = _lib.load().unirec_dropout_add(x.data_ptr(), ldx, x_row_mod, _ptr(residual), ldres, out.data_ptr(), H, rows, H,
                                        thr16, seed, site, off, _stream())
    _lib.check(rc, "unirec_dropout_add")
    return out


def dropout_backward(dy: torch.Tensor, dropout: Tuple[int, int, int]) -> torch.Tensor:
    """dx = dy o mask * scale for the mask of (thr16, seed, site); bf16 [rows, H]."""
    _req(dy, torch.bfloat16, "dropout_backward.dy")
    rows, H, lddy = _rows2d(dy, "dropout_backward.dy")
    dx = torch.empty(rows, H, device=dy.device, dtype=torch.bfloat16)
    thr16, seed, site, off = _drop_args(dropout)
    rc = _lib.load().unirec_dropout_backward(dy.data_ptr(), lddy, dx.data_ptr(), H, rows, H, thr16, seed, site, off,
                                             _stream())
    _lib.check(rc, "unirec_dropout_backward")
    return dx


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    if x.dtype == torch.bfloat16:
        return x
    _req(x, torch.float32, "cast_bf16.x")
    x = x.contiguous()
    out = torch.empty_like(x, dtype=torch.bfloat16)
    rc = _lib.load().unirec_cast_f32_to_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "unirec_cast_f32_to_bf16")
    return out


def mean_tokens(x: torch.Tensor, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """x bf16 [B, T, H] -> [B, H] mean over T."""
    _req(x, torch.bfloat16, "mean_tokens.x")
    if x.dim() != 3 or not x.is_contiguous():
        raise RuntimeError("mean_tokens: x must be contiguous [B, T, H]")
    B, T, H = x.shape
    out = torch.empty(B, H, device=x.device, dtype=out_dtype)
    rc = _lib.load().unirec_mean_tokens(x.data_ptr(), H, B, T, H, out.data_ptr(), H,
                                        1 if out_dtype == torch.float32 else 0, _stream())
    _lib.check(rc, "unirec_mean_tokens")
    return out


def field_projection(rec: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                     out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """rec bf16 [B, T, E], weight fp32 [F, T], bias fp32 [F] -> [B, F, E]."""
    _req(rec, torch.bfloat16, "field_projection.rec")
    _req(weight, torch.float32, "field_projection.weight")
    _req(bias, torch.float32, "field_projection.bias")
    B, T, E = rec.shape
    F = weight.shape[0]
    out = torch.empty(B, F, E, device=rec.device, dtype=out_dtype)
    rc = _lib.load().unirec_field_projection(rec.contiguous().data_ptr(), weight.contiguous().data_ptr(),
                                             bias.data_ptr(), out.data_ptr(), 1 if out_dtype == torch.float32 else 0,
                                             B, T, F, E, _stream())
    _lib.check(rc, "unirec_field_projection")
    return out


_pe_tables = {}


def positional_encoding_table(S: int, D: int, device) -> torch.Tensor:
    """fp32 [S, D] sinusoidal table (models/user_sequence_encoder.py:20-24), built once per (S, D, device) by the
    CUDA kernel and kept resident (6.5 MB at S = 1600, D = 1024)."""
    key = (int(S), int(D), torch.device(device))
    t = _pe_tables.get(key)
    if t is None:
        t = torch.empty(S, D, device=device, dtype=torch.float32)
        rc = _lib.load().unirec_positional_encoding(t.data_ptr(), S, D, _stream())
        _lib.check(rc, "unirec_positional_encoding")
        _pe_tables[key] = t
    return t


def build_user_sequence(table: torch.Tensor, history: torch.Tensor, lengths: torch.Tensor,
                        context: Optional[torch.Tensor] = None, use_pe_table: bool = True
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """table bf16 [N, Q, D]; history int64 [B, Hmax]; lengths int32 [B]; context bf16 [B, Hmax, D] or None.
    Returns (seq bf16 [B, Hmax*Q, D], mask fp32 [B, Hmax*Q])."""
    _req(table, torch.bfloat16, "build_user_sequence.table")
    _req(history, torch.int64, "build_user_sequence.history")
    _req(lengths, torch.int32, "build_user_sequence.lengths")
    if not (table.is_contiguous() and history.is_contiguous() and lengths.is_contiguous()):
        raise RuntimeError("build_user_sequence: inputs must be contiguous")
    N, Q, D = table.shape
    B, Hmax = history.shape
    if context is not None:
        _req(context, torch.bfloat16, "build_user_sequence.context")
        context = context.contiguous()
    seq = torch.empty(B, Hmax * Q, D, device=table.device, dtype=torch.bfloat16)
    mask = torch.empty(B, Hmax * Q, device=table.device, dtype=torch.float32)
    pe = positional_encoding_table(Hmax * Q, D, table.device) if use_pe_table else None
    rc = _lib.load().unirec_build_user_sequence(table.data_ptr(), N, history.data_ptr(), lengths.data_ptr(),
                                                _ptr(context), _ptr(pe), seq.data_ptr(), mask.data_ptr(), B, Hmax, Q, D,
                                                _stream())
    _lib.check(rc, "unirec_build_user_sequence")
    return seq, mask


def inv_l2_norm(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise RuntimeError("inv_l2_norm: x must be bf16 or fp32")
    _req(x, x.dtype, "inv_l2_norm.x")
    rows, D, ldx = _rows2d(x, "inv_l2_norm.x")
    out = torch.empty(rows, device=x.device, dtype=torch.float32)
    rc = _lib.load().unirec_inv_l2_norm(x.data_ptr(), 1 if x.dtype == torch.float32 else 0, ldx, out.data_ptr(), rows,
                                        D, float(eps), _stream())
    _lib.check(rc, "unirec_inv_l2_norm")
    return out


def score_topk(users: torch.Tensor, cands: torch.Tensor, k: int, *, user_inv: Optional[torch.Tensor] = None,
               cand_inv: Optional[torch.Tensor] = None, index_base: int = 0
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Cosine top-k of users bf16 [B, D] against cands bf16 [N, D] -> (scores fp32 [B,k], idx int64 [B,k])."""
    _req(users, torch.bfloat16, "score_topk.users")
    _req(cands, torch.bfloat16, "score_topk.cands")
    B, D = users.shape
    N, D2 = cands.shape
    if D != D2:
        raise RuntimeError("score_topk: dims differ")
    if user_inv is None:
        user_inv = inv_l2_norm(users)
    if cand_inv is None:
        cand_inv = inv_l2_norm(cands)
    lib = _lib.load()
    ws_bytes = int(lib.unirec_score_topk_workspace_bytes(B, N, k))
    if ws_bytes < 0:
        _lib.check(1, "unirec_score_topk_workspace_bytes")
    ws = torch.empty(max(ws_bytes, 16), device=users.device, dtype=torch.uint8)
    scores = torch.empty(B, k, device=users.device, dtype=torch.float32)
    idx = torch.empty(B, k, device=users.device, dtype=torch.int64)
    with _Timed("score_topk", 2.0 * B * N * D):
        rc = lib.unirec_score_topk(users.data_ptr(), users.stride(0), user_inv.data_ptr(), cands.data_ptr(),
                                   cands.stride(0), cand_inv.data_ptr(), B, N, D, k, index_base, scores.data_ptr(),
                                   idx.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
    _lib.check(rc, "unirec_score_topk")
    return scores, idx


def topk_merge(scores: torch.Tensor, idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """scores fp32 [G, B, k], idx int64 [G, B, k] (each list descending) -> merged [B, k]."""
    _req(scores, torch.float32, "topk_merge.scores")
    _req(idx, torch.int64, "topk_merge.idx")
    G, B, k = scores.shape
    scores = scores.contiguous()
    idx = idx.contiguous()
    out_s = torch.empty(B, k, device=scores.device, dtype=torch.float32)
    out_i = torch.empty(B, k, device=scores.device, dtype=torch.int64)
    rc = _lib.load().unirec_topk_merge(scores.data_ptr(), idx.data_ptr(), G, B, k, out_s.data_ptr(), out_i.data_ptr(),
                                       _stream())
    _lib.check(rc, "unirec_topk_merge")
    return out_s, out_i


# ------------------------------------------------------------------------------------------------
# Backward-pass ops of the item Q-Former training step (C ABI: "Backward pass" block of the header)
# ------------------------------------------------------------------------------------------------
def gemm_general(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool, b_mn: bool, M: int, N: int, K: int,
                 out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
                 accumulate: bool = False, ksplit: int = 0) -> torch.Tensor:
    """out[M,N] (+)= sum_k A(m,k) B(n,k).  a is stored [M,K] (a_mn False) or [K,M] (a_mn True); b is stored
    [N,K] (b_mn False) or [K,N] (b_mn True); 2-D row views with contiguous last dim."""
    _req(a, torch.bfloat16, "gemm_general.a")
    _req(b, torch.bfloat16, "gemm_general.b")
    if a.dim() != 2 or b.dim() != 2:
        raise RuntimeError("gemm_general: operands must be 2-D row views")
    if tuple(a.shape) != ((K, M) if a_mn else (M, K)) or tuple(b.shape) != ((K, N) if b_mn else (N, K)):
        raise RuntimeError(f"gemm_general: operand shapes {tuple(a.shape)}, {tuple(b.shape)} do not match M={M} N={N} K={K}")
    if out is None:
        if accumulate:
            raise RuntimeError("gemm_general: accumulate needs an existing fp32 out tensor")
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    if out.dim() != 2 or tuple(out.shape) != (M, N) or out.stride(1) != 1:
        raise RuntimeError("gemm_general: out must be a 2-D [M, N] row view")
    if out.dtype not in (torch.bfloat16, torch.float32) or (accumulate and out.dtype != torch.float32):
        raise RuntimeError("gemm_general: out must be bf16 or fp32 (fp32 when accumulating)")
    with _Timed("gemm", 2.0 * M * N * K):
        rc = _lib.load().unirec_gemm_general(a.data_ptr(), a.stride(0), 1 if a_mn else 0, b.data_ptr(), b.stride(0),
                                             1 if b_mn else 0, out.data_ptr(), out.stride(0),
                                             1 if out.dtype == torch.float32 else 0, 1 if accumulate else 0, M, N, K,
                                             ksplit, _stream())
    _lib.check(rc, "unirec_gemm_general")
    return out


def linear_dgrad(dy: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """dx[M, K_in] = dy[M, N] @ weight[N, K_in]  (weight bf16 as stored by nn.Linear)."""
    M, N = dy.shape
    return gemm_general(dy, weight, a_mn=False, b_mn=True, M=M, N=weight.shape[1], K=N)


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor) -> torch.Tensor:
    """dw[N, K_in] += dy[rows, N]^T @ x[rows, K_in]  (dw fp32, accumulated with atomics; zero it first)."""
    rows, N = dy.shape
    return gemm_general(dy, x, a_mn=True, b_mn=True, M=N, N=x.shape[1], K=rows, out=dw, accumulate=True)


def gelu(z: torch.Tensor) -> torch.Tensor:
    _req(z, torch.bfloat16, "gelu.z")
    z = z.contiguous()
    out = torch.empty_like(z)
    rc = _lib.load().unirec_gelu_forward(z.data_ptr(), out.data_ptr(), z.numel(), _stream())
    _lib.check(rc, "unirec_gelu_forward")
    return out


def gelu_backward(z: torch.Tensor, da: torch.Tensor) -> torch.Tensor:
    _req(z, torch.bfloat16, "gelu_backward.z")
    _req(da, torch.bfloat16, "gelu_backward.da")
    if not (z.is_contiguous() and da.is_contiguous()) or z.shape != da.shape:
        raise RuntimeError("gelu_backward: z and da must be contiguous and of the same shape")
    dz = torch.empty_like(z)
    rc = _lib.load().unirec_gelu_backward(z.data_ptr(), da.data_ptr(), dz.data_ptr(), z.numel(), _stream())
    _lib.check(rc, "unirec_gelu_backward")
    return dz


def colsum(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[n] += sum_rows x[row, n]; x bf16 2-D row view, out fp32 [N]."""
    _req(x, torch.bfloat16, "colsum.x")
    _req(out, torch.float32, "colsum.out")
    rows, N = x.shape
    rc = _lib.load().unirec_colsum(x.data_ptr(), x.stride(0), rows, N, out.data_ptr(), _stream())
    _lib.check(rc, "unirec_colsum")
    return out


def layernorm_backward(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float, dgamma: torch.Tensor,
                       dbeta: torch.Tensor, dy2: Optional[torch.Tensor] = None, *,
                       dropout: Optional[Tuple] = None, dbias: Optional[torch.Tensor] = None):
    """x = LayerNorm input (bf16 [rows, H]), dy (+ dy2) = gradient of its output; returns dx bf16 and accumulates
    dgamma / dbeta (fp32 [H]).  Fused extras for the block LayerNorm(dropout(dense(.)) + input): with
    dropout = (thr16, seed, site[, seed_offset]) the call returns (dx, dx_drop) where dx_drop = dropout_backward(dx);
    dbias (fp32 [H]) accumulates the column sums of dx_drop (of dx without dropout) - the dense layer's bias gradient."""
    for t_, n in ((x, "x"), (dy, "dy")):
        _req(t_, torch.bfloat16, f"layernorm_backward.{n}")
    rows, H, ldx = _rows2d(x, "layernorm_backward.x")
    _, _, lddy = _rows2d(dy, "layernorm_backward.dy")
    lddy2 = 0
    if dy2 is not None:
        _req(dy2, torch.bfloat16, "layernorm_backward.dy2")
        _, _, lddy2 = _rows2d(dy2, "layernorm_backward.dy2")
    dx = torch.empty(rows, H, device=x.device, dtype=torch.bfloat16)
    if dropout is None and dbias is None:
        rc = _lib.load().unirec_layernorm_backward(x.data_ptr(), ldx, dy.data_ptr(), lddy, _ptr(dy2), lddy2,
                                                   gamma.data_ptr(), float(eps), dx.data_ptr(), H, dgamma.data_ptr(),
                                                   dbeta.data_ptr(), rows, H, _stream())
        _lib.check(rc, "unirec_layernorm_backward")
        return dx
    thr16, seed, site, off = _drop_args(dropout)
    dx_drop = torch.empty(rows, H, device=x.device, dtype=torch.bfloat16) if dropout is not None else None
    if dbias is not None:
        _req(dbias, torch.float32, "layernorm_backward.dbias")
    rc = _lib.load().unirec_layernorm_backward_fused(
        x.data_ptr(), ldx, dy.data_ptr(), lddy, _ptr(dy2), lddy2, gamma.data_ptr(), float(eps), dx.data_ptr(), H,
        dgamma.data_ptr(), dbeta.data_ptr(), rows, H, thr16, seed, site, off, _ptr(dx_drop), H, _ptr(dbias), _stream())
    _lib.check(rc, "unirec_layernorm_backward_fused")
    return dx if dropout is None else (dx, dx_drop)


def attention_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, dout: torch.Tensor, dq: torch.Tensor,
                       dk: torch.Tensor, dv: torch.Tensor, *, batch: int, num_heads: int, nq: int, nk: int,
                       key_mask: Optional[torch.Tensor] = None, dropout: Optional[Tuple[int, int, int]] = None):
    """Backward of `attention` for nq, nk <= 64; dq/dk/dv are preallocated bf16 row views (written in place)."""
    for t_, n in ((q, "q"), (k, "k"), (v, "v"), (dout, "dout"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
        _req(t_, torch.bfloat16, f"attention_backward.{n}")
        if t_.dim() != 2:
            raise RuntimeError("attention_backward: tensors must be 2-D row views")
    if key_mask is not None:
        _req(key_mask, torch.float32, "attention_backward.key_mask")
    thr16, seed, site, off = _drop_args(dropout)
    rc = _lib.load().unirec_attention_dropout_backward(
        q.data_ptr(), q.stride(0), nq, k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), nk, _ptr(key_mask),
        dout.data_ptr(), dout.stride(0), dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0), dv.data_ptr(),
        dv.stride(0), batch, num_heads, nq, nk, 64, 0.125, thr16, seed, site, off, _stream())
    _lib.check(rc, "unirec_attention_dropout_backward")


# ------------------------------------------------------------------------------------------------
# Per-user candidate-list scoring and token injection of the joint trainer (C ABI: "candidate-LIST scoring" block)
# ------------------------------------------------------------------------------------------------
def _list_operands(users, pos, cands, mask, offsets, max_list):
    if users.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("list_scores: embeddings must be fp32 or bf16")
    for t, n in ((users, "users"), (pos, "pos"), (cands, "cands")):
        _req(t, users.dtype, f"list_scores.{n}")
    if users.dim() != 2 or pos.shape != users.shape:
        raise RuntimeError("list_scores: users and pos must both be [B, D]")
    B, D = users.shape
    if offsets is not None:
        if mask is not None:
            raise RuntimeError("list_scores: pass mask (padded lists) or offsets (ragged lists), not both")
        _req(offsets, torch.int64, "list_scores.offsets")
        if cands.dim() != 2 or cands.shape[1] != D or offsets.numel() != B + 1 or max_list is None:
            raise RuntimeError("list_scores: ragged lists need cands [total, D], offsets [B + 1] and max_list")
        C, ldc = int(max_list), cands.stride(0)
        offsets = offsets.contiguous()
    else:
        if cands.dim() != 3 or cands.shape[0] != B or cands.shape[2] != D or not cands.is_contiguous():
            raise RuntimeError("list_scores: padded lists need contiguous cands [B, C, D]")
        C, ldc = cands.shape[1], D
        if mask is not None:
            _req(mask, torch.bool, "list_scores.mask")
            if tuple(mask.shape) != (B, C):
                raise RuntimeError("list_scores: mask must be [B, C]")
            mask = mask.contiguous()
    return B, C, D, ldc, mask, offsets


def list_scores(users: torch.Tensor, pos: torch.Tensor, cands: torch.Tensor, *, mask: Optional[torch.Tensor] = None,
                offsets: Optional[torch.Tensor] = None, max_list: Optional[int] = None, eps: float = 1e-12
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Cosine similarity of every user with its own list: entry 0 = pos[b], entry 1 + c = its c-th negative.
    Padded lists: cands [B, C, D] (+ bool mask [B, C], False = padding); ragged lists: cands [total, D], offsets int64
    [B + 1], max_list = longest list.  Returns (sims, inv_norm), both fp32 [B, 1 + C]; padding scores -inf."""
    B, C, D, ldc, mask, offsets = _list_operands(users, pos, cands, mask, offsets, max_list)
    sims = torch.empty(B, C + 1, device=users.device, dtype=torch.float32)
    inv = torch.empty(B, C + 1, device=users.device, dtype=torch.float32)
    es = users.element_size()
    step = 65535                                    # users are the grid's y dimension: larger batches go in slices
    with _Timed("list_scores", float(B) * (C + 2) * D * es):
        for lo in range(0, B, step):
            n = min(step, B - lo)
            # padded lists: the slice's rows start at cands[lo]; ragged lists: offsets are absolute row numbers
            c_ptr = cands.data_ptr() + (lo * C * ldc * es if offsets is None else 0)
            rc = _lib.load().unirec_list_scores(
                users.data_ptr() + lo * users.stride(0) * es, users.stride(0), pos.data_ptr() + lo * pos.stride(0) * es,
                pos.stride(0), c_ptr, ldc, 1 if users.dtype == torch.float32 else 0,
                None if mask is None else mask.data_ptr() + lo * C,
                None if offsets is None else offsets.data_ptr() + lo * 8, n, C, D, float(eps),
                sims.data_ptr() + lo * (C + 1) * 4, inv.data_ptr() + lo * (C + 1) * 4, _stream())
            _lib.check(rc, "unirec_list_scores")
    return sims, inv


def infonce_rank(sims: torch.Tensor, temperature: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """sims fp32 [B, 1 + C] (column 0 = positive, -inf = padding) -> (InfoNCE loss per user fp32 [B], 1-based rank of
    the positive int32 [B])."""
    _req(sims, torch.float32, "infonce_rank.sims")
    if sims.dim() != 2 or not sims.is_contiguous():
        raise RuntimeError("infonce_rank: sims must be contiguous [B, 1 + C]")
    B, C1 = sims.shape
    loss = torch.empty(B, device=sims.device, dtype=torch.float32)
    rank = torch.empty(B, device=sims.device, dtype=torch.int32)
    rc = _lib.load().unirec_infonce_rank(sims.data_ptr(), B, C1 - 1, float(temperature), loss.data_ptr(), rank.data_ptr(),
                                         _stream())
    _lib.check(rc, "unirec_infonce_rank")
    return loss, rank


def list_scores_backward(users, pos, cands, sims, inv_norm, dloss, temperature: float, *, mask=None, offsets=None,
                         max_list=None, eps: float = 1e-12, want_list_grad: bool = False):
    """Gradient of sum_b dloss[b] * infonce_loss[b]: returns (d_user fp32 [B, D], d_list fp32 [B, 1 + C, D] or None)."""
    B, C, D, ldc, mask, offsets = _list_operands(users, pos, cands, mask, offsets, max_list)
    _req(dloss, torch.float32, "list_scores_backward.dloss")
    d_user = torch.zeros(B, D, device=users.device, dtype=torch.float32)
    d_list = torch.empty(B, C + 1, D, device=users.device, dtype=torch.float32) if want_list_grad else None
    rc = _lib.load().unirec_list_scores_backward(
        users.data_ptr(), users.stride(0), pos.data_ptr(), pos.stride(0), cands.data_ptr(), ldc,
        1 if users.dtype == torch.float32 else 0, _ptr(mask), _ptr(offsets), B, C, D, float(eps), sims.data_ptr(),
        inv_norm.data_ptr(), dloss.contiguous().data_ptr(), float(temperature), d_user.data_ptr(), _ptr(d_list), _stream())
    _lib.check(rc, "unirec_list_scores_backward")
    return d_user, d_list


def inject_tokens(text_embeds: torch.Tensor, input_ids: torch.Tensor, token_ids: torch.Tensor,
                  tokens: torch.Tensor) -> torch.Tensor:
    """In place: text_embeds[b, s, :] = tokens[b, slot, :] wherever input_ids[b, s] == token_ids[slot].
    text_embeds [B, S, Hd] (fp32 / bf16, contiguous), input_ids int64 [B, S], token_ids int64 [slots],
    tokens [B, slots, Hd] (fp32 / bf16)."""
    for t, n in ((text_embeds, "text_embeds"), (tokens, "tokens")):
        if t.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError(f"inject_tokens: {n} must be fp32 or bf16")
        _req(t, t.dtype, f"inject_tokens.{n}")
    _req(input_ids, torch.int64, "inject_tokens.input_ids")
    _req(token_ids, torch.int64, "inject_tokens.token_ids")
    B, S, Hd = text_embeds.shape
    slots = token_ids.numel()
    if not text_embeds.is_contiguous() or tuple(input_ids.shape) != (B, S) or tuple(tokens.shape) != (B, slots, Hd):
        raise RuntimeError("inject_tokens: expected contiguous text_embeds [B, S, Hd], input_ids [B, S], tokens [B, slots, Hd]")
    rc = _lib.load().unirec_inject_tokens(input_ids.contiguous().data_ptr(), B, S, token_ids.contiguous().data_ptr(), slots,
                                          tokens.contiguous().data_ptr(), 1 if tokens.dtype == torch.float32 else 0,
                                          text_embeds.data_ptr(), 1 if text_embeds.dtype == torch.float32 else 0, Hd, Hd,
                                          _stream())
    _lib.check(rc, "unirec_inject_tokens")
    return text_embeds


def inject_tokens_backward(d_text: torch.Tensor, input_ids: torch.Tensor, token_ids: torch.Tensor) -> torch.Tensor:
    """Backward of `inject_tokens`: d_text [B, S, Hd] (fp32 / bf16, contiguous) is the upstream gradient and is edited in
    place (rows of overwritten positions -> 0); returns d_tokens fp32 [B, slots, Hd]."""
    if d_text.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("inject_tokens_backward: d_text must be fp32 or bf16")
    _req(d_text, d_text.dtype, "inject_tokens_backward.d_text")
    _req(input_ids, torch.int64, "inject_tokens_backward.input_ids")
    _req(token_ids, torch.int64, "inject_tokens_backward.token_ids")
    B, S, Hd = d_text.shape
    slots = token_ids.numel()
    if not d_text.is_contiguous() or tuple(input_ids.shape) != (B, S):
        raise RuntimeError("inject_tokens_backward: expected contiguous d_text [B, S, Hd] and input_ids [B, S]")
    d_tokens = torch.zeros(B, slots, Hd, device=d_text.device, dtype=torch.float32)
    rc = _lib.load().unirec_inject_tokens_backward(input_ids.contiguous().data_ptr(), B, S, token_ids.contiguous().data_ptr(),
                                                   slots, d_text.data_ptr(), 1 if d_text.dtype == torch.float32 else 0, Hd,
                                                   d_tokens.data_ptr(), Hd, _stream())
    _lib.check(rc, "unirec_inject_tokens_backward")
    return d_tokens


def context_hidden(timestamps: Optional[torch.Tensor], coords: Optional[torch.Tensor], w1t, b1t, w1g, b1g, hidden: int,
                   want_features: bool = False):
    """First halves of the Timestamp / GeoCoordinate encoders (models/mwne.py:504-610) for n events: returns bf16
    [n, 2 * hidden] = [gelu(W1t f_time + b1t) | gelu(W1g f_geo + b1g)] (and the fp32 [n, 12] raw features if asked).
    timestamps [n] fp32 / int64 or None; coords fp32 [n, 2] (lat, lon degrees) or None."""
    if timestamps is None and coords is None:
        raise RuntimeError("context_hidden: timestamps and coords are both None")
    ref = timestamps if timestamps is not None else coords
    if not ref.is_cuda:
        raise RuntimeError("context_hidden: expected CUDA tensors (unirec_b200 has no CPU path)")
    ts_int64 = 0
    if timestamps is not None:
        if timestamps.dtype in (torch.float64, torch.int32):
            timestamps = timestamps.float()          # the reference's `.float()` (models/mwne.py:527)
        if timestamps.dtype not in (torch.float32, torch.int64):
            raise RuntimeError("context_hidden: timestamps must be fp32 / fp64 / int32 / int64")
        ts_int64 = 1 if timestamps.dtype == torch.int64 else 0
        timestamps = timestamps.contiguous().view(-1)
        n = timestamps.numel()
        for t, nm in ((w1t, "w1t"), (b1t, "b1t")):
            _req(t, torch.float32, f"context_hidden.{nm}")
        if tuple(w1t.shape) != (hidden, 9) or not w1t.is_contiguous():
            raise RuntimeError("context_hidden: w1t must be contiguous [hidden, 9]")
    if coords is not None:
        _req(coords, torch.float32, "context_hidden.coords")
        if coords.dim() != 2 or coords.shape[1] != 2:
            raise ValueError("Input coordinates must be of shape [batch_size, 2]")      # models/mwne.py:593-594
        coords = coords.contiguous()
        n = coords.shape[0]
        for t, nm in ((w1g, "w1g"), (b1g, "b1g")):
            _req(t, torch.float32, f"context_hidden.{nm}")
        if tuple(w1g.shape) != (hidden, 3) or not w1g.is_contiguous():
            raise RuntimeError("context_hidden: w1g must be contiguous [hidden, 3]")
    if timestamps is not None and coords is not None and timestamps.numel() != coords.shape[0]:
        raise RuntimeError("context_hidden: timestamps and coords describe different numbers of events")
    out = torch.empty(n, 2 * hidden, device=ref.device, dtype=torch.bfloat16)
    feats = torch.zeros(n, 12, device=ref.device, dtype=torch.float32) if want_features else None
    rc = _lib.load().unirec_context_hidden(_ptr(timestamps), ts_int64, _ptr(coords), _ptr(w1t) if timestamps is not None else None,
                                           _ptr(b1t) if timestamps is not None else None,
                                           _ptr(w1g) if coords is not None else None,
                                           _ptr(b1g) if coords is not None else None, n, hidden, out.data_ptr(), 2 * hidden,
                                           _ptr(feats), _stream())
    _lib.check(rc, "unirec_context_hidden")
    return (out, feats) if want_features else out


def mwne_encode(numbers: torch.Tensor, freqs: torch.Tensor, fourier_w: torch.Tensor, raw_scale: Optional[torch.Tensor],
                extra_w: Optional[torch.Tensor], scale: Optional[torch.Tensor], dim: int,
                out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """ImprovedMathematicalEncoder.forward (+ optional eval-mode normaliser scale); numbers fp32 [n] -> [n, dim]."""
    _req(numbers, torch.float32, "mwne_encode.numbers")
    for t, nm in ((freqs, "freqs"), (fourier_w, "fourier_w"), (raw_scale, "raw_scale"), (extra_w, "extra_w"),
                  (scale, "scale")):
        if t is not None:
            _req(t, torch.float32, f"mwne_encode.{nm}")
            if not t.is_contiguous():
                raise RuntimeError(f"mwne_encode.{nm} must be contiguous")
    x = numbers.contiguous().view(-1)
    n = x.numel()
    out = torch.empty(n, dim, device=numbers.device, dtype=out_dtype)
    rc = _lib.load().unirec_mwne_encode(x.data_ptr(), n, freqs.data_ptr(), freqs.numel(), fourier_w.data_ptr(),
                                        _ptr(raw_scale), _ptr(extra_w), _ptr(scale), dim, out.data_ptr(),
                                        1 if out_dtype == torch.float32 else 0, _stream())
    _lib.check(rc, "unirec_mwne_encode")
    return out


def reconstruction_metrics(reconstructed: torch.Tensor, original: torch.Tensor, attention_mask: torch.Tensor,
                           acc: Optional[torch.Tensor] = None, eps: float = 1e-12) -> torch.Tensor:
    """acc (float64 [3], created zeroed if None) += [sum of squared errors, sum of cosine similarities, count] over the
    (item, field) rows with attention_mask != 0.  reconstructed [B, F, E] fp32 / bf16, original fp32 [B, F, E]."""
    if reconstructed.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("reconstruction_metrics: reconstructed must be fp32 or bf16")
    _req(reconstructed, reconstructed.dtype, "reconstruction_metrics.reconstructed")
    _req(original, torch.float32, "reconstruction_metrics.original")
    if reconstructed.shape != original.shape or reconstructed.dim() != 3:
        raise RuntimeError("reconstruction_metrics: expected two [B, F, E] tensors")
    B, F_, E = reconstructed.shape
    if not attention_mask.is_cuda or tuple(attention_mask.shape) != (B, F_):
        raise RuntimeError("reconstruction_metrics: attention_mask must be a CUDA tensor [B, F]")
    m = attention_mask.to(torch.float32).contiguous()
    if acc is None:
        acc = torch.zeros(3, device=reconstructed.device, dtype=torch.float64)
    _req(acc, torch.float64, "reconstruction_metrics.acc")
    rc = _lib.load().unirec_reconstruction_metrics(reconstructed.contiguous().data_ptr(),
                                                   1 if reconstructed.dtype == torch.float32 else 0,
                                                   original.contiguous().data_ptr(), m.data_ptr(), B * F_, E, float(eps),
                                                   acc.data_ptr(), _stream())
    _lib.check(rc, "unirec_reconstruction_metrics")
    return acc


def linear_gather(table: torch.Tensor, history: torch.Tensor, lengths: torch.Tensor, pad_table: torch.Tensor,
                  weight: torch.Tensor, bias: Optional[torch.Tensor], posbias: torch.Tensor) -> torch.Tensor:
    """K/V projection of the (never materialised) user sequences: table bf16 [N_items, 32, K]; history int64 [B, Hmax];
    lengths int32 [B]; pad_table bf16 [Hmax * 32, K] (= -PE); weight bf16 [N, K]; bias fp32 [N]; posbias bf16
    [Hmax * 32 + 128, N] (= PE W^T, first 128 rows repeated at the end).  Returns bf16 [B * Hmax * 32, N]."""
    _req(table, torch.bfloat16, "linear_gather.table")
    _req(history, torch.int64, "linear_gather.history")
    _req(lengths, torch.int32, "linear_gather.lengths")
    _req(pad_table, torch.bfloat16, "linear_gather.pad_table")
    _req(weight, torch.bfloat16, "linear_gather.weight")
    _req(posbias, torch.bfloat16, "linear_gather.posbias")
    if bias is not None:
        _req(bias, torch.float32, "linear_gather.bias")
    if table.dim() != 3 or table.shape[1] != 32 or not table.is_contiguous():
        raise RuntimeError("linear_gather: table must be contiguous [N_items, 32, K]")
    n_items, _, K = table.shape
    B, Hmax = history.shape
    N = weight.shape[0]
    period = Hmax * 32
    if not (history.is_contiguous() and lengths.is_contiguous()) or lengths.numel() != B:
        raise RuntimeError("linear_gather: history [B, Hmax] and lengths [B] must be contiguous")
    if tuple(pad_table.shape) != (period, K) or posbias.shape[0] < period + 128 or posbias.shape[1] != N:
        raise RuntimeError("linear_gather: pad_table must be [Hmax * 32, K] and posbias [>= Hmax * 32 + 128, N]")
    M = B * period
    out = torch.empty(M, N, device=table.device, dtype=torch.bfloat16)
    with _Timed("gemm", 2.0 * M * N * K, f"{M}x{N}x{K}"):
        rc = _lib.load().unirec_linear_gather_bf16(table.data_ptr(), K, n_items * 32, history.data_ptr(), lengths.data_ptr(),
                                                   Hmax, pad_table.data_ptr(), pad_table.stride(0), period, weight.data_ptr(),
                                                   weight.stride(0), _ptr(bias), posbias.data_ptr(), posbias.stride(0),
                                                   posbias.shape[0], period, out.data_ptr(), out.stride(0), M, N, K, _stream())
    _lib.check(rc, "unirec_linear_gather_bf16")
    return out
