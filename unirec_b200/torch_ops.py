"""torch custom-op layer over the C ABI: `torch.ops.unirec_b200.*`.

The reference has no custom ops of its own; this layer is what BASELINE.json's north_star calls the
"thin C-ABI torch custom-op layer": every forward kernel of the path is registered with `torch.library.custom_op`
(CUDA only - there is no CPU kernel to dispatch to, a CPU tensor raises NotImplementedError) together with a fake /
meta implementation for shape inference, so the ops can be used under FakeTensorMode / `torch.export` and from code
that only knows the dispatcher.  Each op body is the ctypes wrapper of `unirec_b200.ops` (one kernel launch on the
current stream).  `linear`, `layernorm` and `attention` also carry `register_autograd` backward formulas on the C ABI's
backward kernels, so a model written against `torch.ops.unirec_b200.*` trains under eager autograd.  The nn.Module
mirrors in `modules.py` call the ctypes wrappers directly - the same launches without the dispatcher's per-call cost
(~10 us x ~435 launches per training step), with the whole backbone as ONE autograd.Function;
`tests/test_kernels_gpu.py::test_torch_ops_match_direct_wrappers` / `::test_torch_ops_autograd` pin the two together.

    import unirec_b200.torch_ops                      # registers the namespace
    y = torch.ops.unirec_b200.linear(x, w, b, None, 0, False)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops

_BF16, _F32 = torch.bfloat16, torch.float32
NAMESPACE = "unirec_b200"


def _dt(fp32: bool):
    return _F32 if fp32 else _BF16


# ------------------------------------------------------------------------------------------------ projections
@torch.library.custom_op(f"{NAMESPACE}::linear", mutates_args=(), device_types="cuda")
def linear(a: Tensor, weight: Tensor, bias: Optional[Tensor], residual: Optional[Tensor], epilogue: int,
           out_fp32: bool) -> Tensor:
    """epilogue(a @ weight.T + bias): 0 bias, 1 bias + erf-GELU, 2 bias + residual (include/unirec_b200.h)."""
    return ops.linear(a, weight, bias, epilogue=epilogue, residual=residual, out_dtype=_dt(out_fp32))


@linear.register_fake
def _(a, weight, bias, residual, epilogue, out_fp32):
    return a.new_empty(*a.shape[:-1], weight.shape[0], dtype=_dt(out_fp32))


@torch.library.custom_op(f"{NAMESPACE}::linear_ln", mutates_args=(), device_types="cuda")
def linear_ln(a: Tensor, weight: Tensor, bias: Tensor, residual: Optional[Tensor], epilogue: int,
              ln_in_stats: Optional[Tensor], ln_in_c: Optional[Tensor], ln_res_stats: Optional[Tensor],
              ln_res_gamma: Optional[Tensor], ln_res_beta: Optional[Tensor], want_stats: bool, eps: float,
              hidden: int) -> Tuple[Tensor, Tensor]:
    """nn.Linear with the LayerNorms around it folded in (unirec_linear_ln_bf16): returns (out bf16, stats fp32
    [M, parts, 2] of out's rows - empty unless want_stats).  ln_in_* : `a` is a LayerNorm input, weight / bias are folded
    (ops.fold_layernorm_weights); ln_res_*: `residual` is a LayerNorm input and enters normalised.  Inference only."""
    M = a.numel() // a.shape[-1]
    stats = ops.ln_stats_buffer(M, weight.shape[0], a.device) if want_stats else a.new_empty(0, dtype=_F32)
    out = ops.linear_ln(a, weight, bias, epilogue=epilogue, residual=residual,
                        ln_in=None if ln_in_stats is None else (ln_in_stats, ln_in_c),
                        ln_res=None if ln_res_stats is None else (ln_res_stats, ln_res_gamma, ln_res_beta),
                        stats_out=stats if want_stats else None, eps=eps, hidden=hidden)
    return out, stats


@linear_ln.register_fake
def _(a, weight, bias, residual, epilogue, ln_in_stats, ln_in_c, ln_res_stats, ln_res_gamma, ln_res_beta, want_stats, eps,
      hidden):
    N = weight.shape[0]
    M = a.numel() // a.shape[-1]
    stats = a.new_empty((M, 2 * (N // 256), 2) if want_stats else (0,), dtype=_F32)
    return a.new_empty(*a.shape[:-1], N, dtype=_BF16), stats


@torch.library.custom_op(f"{NAMESPACE}::layernorm", mutates_args=(), device_types="cuda")
def layernorm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, residual: Optional[Tensor], rows: int,
              in_row_mod: int, out_fp32: bool) -> Tensor:
    """LayerNorm(x [+ residual]); rows = 0: as many rows as x; in_row_mod > 0: x rows broadcast over `rows` rows."""
    return ops.layernorm(x, gamma, beta, eps, residual=residual, rows=rows if rows > 0 else None,
                         in_row_mod=in_row_mod, out_dtype=_dt(out_fp32))


@layernorm.register_fake
def _(x, gamma, beta, eps, residual, rows, in_row_mod, out_fp32):
    H = x.shape[-1]
    if in_row_mod > 0 or x.dim() == 2:
        return x.new_empty(rows if rows > 0 else x.numel() // H, H, dtype=_dt(out_fp32))
    return x.new_empty(x.shape, dtype=_dt(out_fp32))


@torch.library.custom_op(f"{NAMESPACE}::attention", mutates_args=(), device_types="cuda")
def attention(q: Tensor, k: Tensor, v: Tensor, key_mask: Optional[Tensor], batch: int, num_heads: int, nq: int,
              nk: int, q_broadcast: bool) -> Tensor:
    """softmax(q k^T / 8 + mask) v per head (head_dim 64), heads merged: bf16 [batch * nq, num_heads * 64]."""
    return ops.attention(q, k, v, batch=batch, num_heads=num_heads, nq=nq, nk=nk, key_mask=key_mask,
                         q_broadcast=q_broadcast)


@attention.register_fake
def _(q, k, v, key_mask, batch, num_heads, nq, nk, q_broadcast):
    return q.new_empty(batch * nq, num_heads * 64, dtype=_BF16)


# ------------------------------------------------------------------------------------------------ autograd
# Backward formulas of the three differentiable building blocks, on the backward kernels of the C ABI (the same ones
# training.BackboneTrainFn drives): tcgen05 dgrad / split-K wgrad GEMMs that read operands in place, the LayerNorm and
# small-tile attention backward kernels.  Eager autograd only (the backward bodies launch kernels through ctypes).
def _linear_setup(ctx, inputs, output):
    a, weight, bias, residual, epilogue, out_fp32 = inputs
    ctx.save_for_backward(a, weight, bias)
    ctx.epilogue, ctx.has_bias, ctx.has_res = epilogue, bias is not None, residual is not None
    ctx.res_dtype = None if residual is None else residual.dtype


def _linear_backward(ctx, dy):
    a, weight, bias = ctx.saved_tensors
    N, K = weight.shape
    dy2 = dy.reshape(-1, N).to(_BF16).contiguous()
    a2 = a.reshape(-1, K)
    if ctx.epilogue == ops.EPI_BIAS_GELU:
        # the fused forward does not keep the pre-activation: recompute it (one GEMM) and apply gelu'(z)
        dy2 = ops.gelu_backward(ops.linear(a2, weight, bias), dy2)
    da = dw = db = dres = None
    if ctx.needs_input_grad[0]:
        da = ops.linear_dgrad(dy2, weight).view(a.shape)
    if ctx.needs_input_grad[1]:
        acc = torch.zeros(N, K, device=dy.device, dtype=_F32)
        rows = dy2.shape[0]
        if rows % 8:
            # the wgrad GEMM contracts over rows in 16-byte steps: zero rows add nothing to dW
            pad = 8 - rows % 8
            ops.linear_wgrad(torch.nn.functional.pad(dy2, (0, 0, 0, pad)), torch.nn.functional.pad(a2, (0, 0, 0, pad)), acc)
        else:
            ops.linear_wgrad(dy2, a2, acc)
        dw = acc.to(weight.dtype)
    if ctx.has_bias and ctx.needs_input_grad[2]:
        db = ops.colsum(dy2, torch.zeros(N, device=dy.device, dtype=_F32))
    if ctx.has_res and ctx.needs_input_grad[3]:
        dres = dy.to(ctx.res_dtype)
    return da, dw, db, dres, None, None


torch.library.register_autograd(f"{NAMESPACE}::linear", _linear_backward, setup_context=_linear_setup)


def _layernorm_setup(ctx, inputs, output):
    x, gamma, beta, eps, residual, rows, in_row_mod, out_fp32 = inputs
    if in_row_mod > 0:
        raise NotImplementedError("unirec_b200::layernorm: autograd of the row-broadcast form is not registered")
    ctx.save_for_backward(x, gamma, residual)
    ctx.eps, ctx.x_dtype = eps, x.dtype


def _layernorm_backward(ctx, dy):
    x, gamma, residual = ctx.saved_tensors
    H = x.shape[-1]
    pre = x.reshape(-1, H)
    if residual is not None:
        pre = (pre.float() + residual.reshape(-1, H).float())
    pre = pre.to(_BF16).contiguous()
    dgamma = torch.zeros(H, device=dy.device, dtype=_F32)
    dbeta = torch.zeros(H, device=dy.device, dtype=_F32)
    dx = ops.layernorm_backward(pre, dy.reshape(-1, H).to(_BF16).contiguous(), gamma, ctx.eps, dgamma, dbeta)
    dres = dx.view(residual.shape) if (residual is not None and ctx.needs_input_grad[4]) else None
    return dx.view(x.shape).to(ctx.x_dtype), dgamma, dbeta, None, dres, None, None, None


torch.library.register_autograd(f"{NAMESPACE}::layernorm", _layernorm_backward, setup_context=_layernorm_setup)


def _attention_setup(ctx, inputs, output):
    q, k, v, key_mask, batch, num_heads, nq, nk, q_broadcast = inputs
    if q_broadcast or nq > 64 or nk > 64:
        raise NotImplementedError("unirec_b200::attention: autograd is registered for per-batch queries and nq, nk <= 64 "
                                  "(the item Q-Former's shapes; the training path of SURVEY.md 8e cfg 2)")
    ctx.save_for_backward(q, k, v, key_mask)
    ctx.dims = (batch, num_heads, nq, nk)


def _attention_backward(ctx, dout):
    q, k, v, key_mask = ctx.saved_tensors
    batch, num_heads, nq, nk = ctx.dims
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops.attention_backward(q, k, v, dout.to(_BF16).contiguous(), dq, dk, dv, batch=batch, num_heads=num_heads, nq=nq,
                           nk=nk, key_mask=key_mask)
    return dq, dk, dv, None, None, None, None, None, None


torch.library.register_autograd(f"{NAMESPACE}::attention", _attention_backward, setup_context=_attention_setup)


# ------------------------------------------------------------------------------------------------ row-wise
@torch.library.custom_op(f"{NAMESPACE}::cast_bf16", mutates_args=(), device_types="cuda")
def cast_bf16(x: Tensor) -> Tensor:
    out = ops.cast_bf16(x)
    return out.clone() if out is x else out          # custom ops may not return an alias of an input


@cast_bf16.register_fake
def _(x):
    return x.new_empty(x.shape, dtype=_BF16)


@torch.library.custom_op(f"{NAMESPACE}::mean_tokens", mutates_args=(), device_types="cuda")
def mean_tokens(x: Tensor, out_fp32: bool) -> Tensor:
    return ops.mean_tokens(x, out_dtype=_dt(out_fp32))


@mean_tokens.register_fake
def _(x, out_fp32):
    return x.new_empty(x.shape[0], x.shape[2], dtype=_dt(out_fp32))


@torch.library.custom_op(f"{NAMESPACE}::field_projection", mutates_args=(), device_types="cuda")
def field_projection(rec: Tensor, weight: Tensor, bias: Tensor, out_fp32: bool) -> Tensor:
    return ops.field_projection(rec, weight, bias, out_dtype=_dt(out_fp32))


@field_projection.register_fake
def _(rec, weight, bias, out_fp32):
    return rec.new_empty(rec.shape[0], weight.shape[0], rec.shape[2], dtype=_dt(out_fp32))


@torch.library.custom_op(f"{NAMESPACE}::build_user_sequence", mutates_args=(), device_types="cuda")
def build_user_sequence(table: Tensor, history: Tensor, lengths: Tensor, context: Optional[Tensor]
                        ) -> Tuple[Tensor, Tensor]:
    return ops.build_user_sequence(table, history, lengths, context)


@build_user_sequence.register_fake
def _(table, history, lengths, context):
    B, Hmax = history.shape
    Q, D = table.shape[1], table.shape[2]
    return table.new_empty(B, Hmax * Q, D, dtype=_BF16), table.new_empty(B, Hmax * Q, dtype=_F32)


@torch.library.custom_op(f"{NAMESPACE}::inv_l2_norm", mutates_args=(), device_types="cuda")
def inv_l2_norm(x: Tensor, eps: float) -> Tensor:
    return ops.inv_l2_norm(x, eps)


@inv_l2_norm.register_fake
def _(x, eps):
    return x.new_empty(x.numel() // x.shape[-1], dtype=_F32)


# ------------------------------------------------------------------------------------------------ ranking
@torch.library.custom_op(f"{NAMESPACE}::score_topk", mutates_args=(), device_types="cuda")
def score_topk(users: Tensor, cands: Tensor, k: int, user_inv: Optional[Tensor], cand_inv: Optional[Tensor],
               index_base: int) -> Tuple[Tensor, Tensor]:
    """Cosine top-k of users [B, D] against cands [N, D]: (scores fp32 [B, k], global indices int64 [B, k])."""
    return ops.score_topk(users, cands, k, user_inv=user_inv, cand_inv=cand_inv, index_base=index_base)


@score_topk.register_fake
def _(users, cands, k, user_inv, cand_inv, index_base):
    B = users.shape[0]
    return users.new_empty(B, k, dtype=_F32), users.new_empty(B, k, dtype=torch.int64)


@torch.library.custom_op(f"{NAMESPACE}::topk_merge", mutates_args=(), device_types="cuda")
def topk_merge(scores: Tensor, idx: Tensor) -> Tuple[Tensor, Tensor]:
    return ops.topk_merge(scores, idx)


@topk_merge.register_fake
def _(scores, idx):
    return scores.new_empty(scores.shape[1], scores.shape[2]), idx.new_empty(idx.shape[1], idx.shape[2])


@torch.library.custom_op(f"{NAMESPACE}::list_scores", mutates_args=(), device_types="cuda")
def list_scores(users: Tensor, pos: Tensor, cands: Tensor, mask: Optional[Tensor], offsets: Optional[Tensor],
                max_list: int, eps: float) -> Tuple[Tensor, Tensor]:
    """Per-user candidate lists (padded + mask, or ragged + offsets): (sims, inv_norm) fp32 [B, 1 + C]."""
    return ops.list_scores(users, pos, cands, mask=mask, offsets=offsets,
                           max_list=max_list if offsets is not None else None, eps=eps)


@list_scores.register_fake
def _(users, pos, cands, mask, offsets, max_list, eps):
    C = max_list if offsets is not None else cands.shape[1]
    return users.new_empty(users.shape[0], C + 1, dtype=_F32), users.new_empty(users.shape[0], C + 1, dtype=_F32)


@torch.library.custom_op(f"{NAMESPACE}::infonce_rank", mutates_args=(), device_types="cuda")
def infonce_rank(sims: Tensor, temperature: float) -> Tuple[Tensor, Tensor]:
    return ops.infonce_rank(sims, temperature)


@infonce_rank.register_fake
def _(sims, temperature):
    return sims.new_empty(sims.shape[0]), sims.new_empty(sims.shape[0], dtype=torch.int32)


OPS = ("linear", "linear_ln", "layernorm", "attention", "cast_bf16", "mean_tokens", "field_projection", "build_user_sequence",
       "inv_l2_norm", "score_topk", "topk_merge", "list_scores", "infonce_rank")
