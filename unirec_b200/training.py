"""Training path of the item Q-Former (BASELINE config 2: forward + backward, data-parallel gradient all-reduce).

Reference step (training/item_qformer_training.py:117-131): out = model(fields, mask); loss = QFormerLoss(...);
loss.backward(); optimizer.step().  Here the backbone (99.4 % of the FLOPs) is ONE torch.autograd.Function whose
forward runs the fused sm_100a kernels and keeps the bf16 activations, and whose backward is written out by hand
on the backward kernels of the C ABI: tcgen05 dgrad / wgrad GEMMs that read nn.Linear weights and activations in
place (MN-major operands, no transposes), fp32 split-K accumulation of weight gradients, LayerNorm / erf-GELU /
small-tile attention backward kernels.  Heads are `LinearFn` (same GEMM kernels).  Loss, optimizer (AdamW) and the
NCCL all-reduce stay in PyTorch.

Dropout (train mode, models/qformer.py:107, :258, :287, :373): masks are a pure function of (seed, site, element)
through Philox4x32-10 (csrc/dropout.cuh), regenerated - never stored - in the backward pass; one fresh seed per forward
call.  The oracle reproduces the same masks on the CPU (oracle/dropout_masks.py), which is what makes a dropped
forward / backward comparable at all.  Probability dropout runs inside the attention kernels; the hidden-state sites
are `dropout_add` (dropout(dense) + residual, one streaming pass) in front of the LayerNorm.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from . import ops


def _f32(t):
    return t.detach().float().contiguous()


class LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM; x bf16 [M, K], W / b fp32 parameters; y bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        w16 = weight.detach().to(torch.bfloat16).contiguous()
        ctx.save_for_backward(x, w16)
        ctx.has_bias = bias is not None
        return ops.linear(x, w16, None if bias is None else _f32(bias))

    @staticmethod
    def backward(ctx, dy):
        x, w16 = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.linear_dgrad(dy, w16) if ctx.needs_input_grad[0] else None
        dw = torch.zeros(w16.shape, device=dy.device, dtype=torch.float32)
        ops.linear_wgrad(dy, x, dw)
        db = None
        if ctx.has_bias:
            db = torch.zeros(w16.shape[0], device=dy.device, dtype=torch.float32)
            ops.colsum(dy, db)
        return dx, dw, db


class BackboneTrainFn(torch.autograd.Function):
    """last_hidden_state = BertModel(query_embeds=queries, encoder_hidden_states=enc, ...) with hand-written
    backward (models/qformer.py:804-972 under autograd).  Inputs after `mask` are the live parameters in the order
    of QFormerBackbone._live_params(); gradients are returned in the same order."""

    SITE_EMBEDDINGS = 0
    KIND_SELF_PROBS, KIND_SELF_OUT, KIND_CROSS_PROBS, KIND_CROSS_OUT, KIND_FFN_OUT = range(5)

    @staticmethod
    def forward(ctx, backbone, drop, enc, mask, query_embeddings, *params):
        """drop = None (dropout off) or (thr16, seed[, seed_offset]); seed_offset = int64 CUDA tensor with one element
        that the kernels add to the seed when they run (CUDA-graph replays, see TrainStepGraph)."""
        cfg = backbone.config

        def site(layer, kind):
            return None if drop is None else (drop[0], drop[1], 1 + 8 * layer + kind) + tuple(drop[2:])

        def dense_residual(x, w, b, res, st):
            # dropout(dense(x)) + res: fused into the GEMM epilogue when there is no dropout
            if st is None:
                return ops.linear(x, w, b, epilogue=ops.EPI_BIAS_RESIDUAL, residual=res)
            return ops.dropout_add(ops.linear(x, w, b), res, st)

        B, S, E = enc.shape
        H, heads = cfg.hidden_size, cfg.num_attention_heads
        Q = query_embeddings.shape[1]
        pk = backbone.packed()
        enc2 = ops.cast_bf16(enc.contiguous()).view(B * S, E)
        m = None if mask is None else mask.to(device=enc.device, dtype=torch.float32).contiguous()
        kv_all = ops.linear(enc2, pk["w_kv_all"], pk["b_kv_all"]) if pk["w_kv_all"] is not None else None
        q0 = query_embeddings.detach().reshape(Q, H).float().contiguous()
        if drop is None:
            h = ops.layernorm(q0, pk["emb_g"], pk["emb_b"], cfg.layer_norm_eps, rows=B * Q, in_row_mod=Q)
        else:
            h = ops.dropout_add(ops.layernorm(q0, pk["emb_g"], pk["emb_b"], cfg.layer_norm_eps), None,
                                (drop[0], drop[1], BackboneTrainFn.SITE_EMBEDDINGS) + tuple(drop[2:]), rows=B * Q, x_row_mod=Q)
        K = BackboneTrainFn
        tape = []
        for li, L in enumerate(pk["layers"]):
            t = {"h_in": h}
            qkv = ops.linear(h, L["w_qkv"], L["b_qkv"])
            ctxt = ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], batch=B, num_heads=heads, nq=Q, nk=Q,
                                 dropout=site(li, K.KIND_SELF_PROBS))
            pre1 = dense_residual(ctxt, L["w_o"], L["b_o"], h, site(li, K.KIND_SELF_OUT))
            h = ops.layernorm(pre1, L["ln1_g"], L["ln1_b"], cfg.layer_norm_eps)
            t.update(qkv=qkv, ctx=ctxt, pre1=pre1, h1=h)
            if L["cross"]:
                qc = ops.linear(h, L["w_qc"], L["b_qc"])
                off = L["kv_slot"] * 2 * H
                ctxc = ops.attention(qc, kv_all[:, off:off + H], kv_all[:, off + H:off + 2 * H], batch=B, num_heads=heads,
                                     nq=Q, nk=S, key_mask=m, dropout=site(li, K.KIND_CROSS_PROBS))
                pre2 = dense_residual(ctxc, L["w_oc"], L["b_oc"], h, site(li, K.KIND_CROSS_OUT))
                h = ops.layernorm(pre2, L["ln2_g"], L["ln2_b"], cfg.layer_norm_eps)
                t.update(qc=qc, ctxc=ctxc, pre2=pre2, h2=h)
            z = ops.linear(h, L["w_1"], L["b_1"])
            a = ops.gelu(z)
            pre3 = dense_residual(a, L["w_2"], L["b_2"], h, site(li, K.KIND_FFN_OUT))
            h = ops.layernorm(pre3, L["ln3_g"], L["ln3_b"], cfg.layer_norm_eps)
            t.update(z=z, a=a, pre3=pre3)
            tape.append(t)
        ctx.backbone, ctx.tape, ctx.kv_all, ctx.enc2, ctx.mask = backbone, tape, kv_all, enc2, m
        ctx.dims = (B, S, Q, H, heads)
        ctx.q0 = q0
        ctx.drop = drop
        return h.view(B, Q, H).float()

    @staticmethod
    def backward(ctx, d_out):
        bb = ctx.backbone
        cfg = bb.config
        pk = bb.packed()
        B, S, Q, H, heads = ctx.dims
        I = cfg.intermediate_size
        dev = d_out.device
        eps = cfg.layer_norm_eps
        M = B * Q
        on_ready: Optional[Callable[[int, List[torch.Tensor]], None]] = getattr(bb, "grad_ready_hook", None)
        on_finish: Optional[Callable[[], None]] = getattr(bb, "grad_finish_hook", None)
        drop = ctx.drop
        K = BackboneTrainFn

        def site(layer, kind):
            return None if drop is None else (drop[0], drop[1], 1 + 8 * layer + kind) + tuple(drop[2:])

        # every parameter gradient of the backbone lives in ONE zero-initialised fp32 buffer (one memset instead of ~190
        # torch.zeros launches per step - the host enqueues this step almost as slowly as the GPU runs it), laid out layer
        # by layer from the last layer to the first, so that a layer's gradients are one contiguous bucket that the
        # data-parallel all-reduce can take in place
        def layer_numel(L):
            n = 3 * H * H + 3 * H + H * H + H + 2 * H + I * H + I + H * I + H + 2 * H
            if L["cross"]:
                n += 2 * (H * H + H) + 2 * H
            return n
        n_cross = sum(1 for L in pk["layers"] if L["cross"])
        E = ctx.enc2.shape[1]
        kv_numel = (2 * H * n_cross) * E + 2 * H * n_cross if n_cross else 0
        flat = torch.zeros(sum(layer_numel(L) for L in pk["layers"]) + kv_numel, device=dev, dtype=torch.float32)
        cursor = [0]

        def zeros(*shape):
            n = 1
            for d_ in shape:
                n *= d_
            v = flat[cursor[0]:cursor[0] + n].view(*shape)
            cursor[0] += n
            return v

        def lin_bwd(dy, x, w16, need_dx=True):
            dw = zeros(*w16.shape)
            ops.linear_wgrad(dy, x, dw)
            db = zeros(w16.shape[0])
            ops.colsum(dy, db)
            return (ops.linear_dgrad(dy, w16) if need_dx else None), dw, db

        def ln_dense_bwd(pre, dy, dy2, gamma, dgamma, dbeta, st, x_act, w16):
            """Backward of LayerNorm(dropout(dense(x_act)) + residual): ONE kernel gives the gradient of the LayerNorm
            input (= the residual branch's gradient), its dropped copy (= the dense output's gradient) and the dense
            bias gradient; then the dense layer's wgrad / dgrad.  Returns (dpre, dx_act, dw, db)."""
            db = zeros(w16.shape[0])
            if st is None:
                dpre = ops.layernorm_backward(pre, dy, gamma, eps, dgamma, dbeta, dy2=dy2, dbias=db)
                ddense = dpre
            else:
                dpre, ddense = ops.layernorm_backward(pre, dy, gamma, eps, dgamma, dbeta, dy2=dy2, dropout=st, dbias=db)
            dw = zeros(*w16.shape)
            ops.linear_wgrad(ddense, x_act, dw)
            return dpre, ops.linear_dgrad(ddense, w16), dw, db

        d_kv_all = torch.zeros(B * S, 2 * H * n_cross, device=dev, dtype=torch.bfloat16) if n_cross else None
        dy = d_out.reshape(M, H).to(torch.bfloat16).contiguous()
        dy2 = None
        layer_grads = [None] * len(pk["layers"])
        for li in range(len(pk["layers"]) - 1, -1, -1):
            L, t = pk["layers"][li], ctx.tape[li]
            g = {}
            bucket_lo = cursor[0]
            # ---- query FFN: h_out = LN3(a W2^T + b2 + h_mid), a = gelu(h_mid W1^T + b1)
            g["ln3_g"], g["ln3_b"] = zeros(H), zeros(H)
            dpre3, da, g["w_2"], g["b_2"] = ln_dense_bwd(t["pre3"], dy, dy2, L["ln3_g"], g["ln3_g"], g["ln3_b"],
                                                         site(li, K.KIND_FFN_OUT), t["a"], L["w_2"])
            dz = ops.gelu_backward(t["z"], da)
            h_mid = t["h2"] if L["cross"] else t["h1"]
            dh_mid, g["w_1"], g["b_1"] = lin_bwd(dz, h_mid, L["w_1"])
            dy, dy2 = dpre3, dh_mid                       # gradient of h_mid = residual branch + FFN branch
            if L["cross"]:
                g["ln2_g"], g["ln2_b"] = zeros(H), zeros(H)
                dpre2, dctxc, g["w_oc"], g["b_oc"] = ln_dense_bwd(t["pre2"], dy, dy2, L["ln2_g"], g["ln2_g"], g["ln2_b"],
                                                                   site(li, K.KIND_CROSS_OUT), t["ctxc"], L["w_oc"])
                off = L["kv_slot"] * 2 * H
                dqc = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
                ops.attention_backward(t["qc"], ctx.kv_all[:, off:off + H], ctx.kv_all[:, off + H:off + 2 * H], dctxc, dqc,
                                       d_kv_all[:, off:off + H], d_kv_all[:, off + H:off + 2 * H], batch=B, num_heads=heads,
                                       nq=Q, nk=S, key_mask=ctx.mask, dropout=site(li, K.KIND_CROSS_PROBS))
                dh1, g["w_qc"], g["b_qc"] = lin_bwd(dqc, t["h1"], L["w_qc"])
                dy, dy2 = dpre2, dh1
            # ---- self-attention block: h1 = LN1(ctx Wo^T + bo + h_in)
            g["ln1_g"], g["ln1_b"] = zeros(H), zeros(H)
            dpre1, dctx, g["w_o"], g["b_o"] = ln_dense_bwd(t["pre1"], dy, dy2, L["ln1_g"], g["ln1_g"], g["ln1_b"],
                                                            site(li, K.KIND_SELF_OUT), t["ctx"], L["w_o"])
            qkv = t["qkv"]
            dqkv = torch.empty(M, 3 * H, device=dev, dtype=torch.bfloat16)
            ops.attention_backward(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], dctx, dqkv[:, :H], dqkv[:, H:2 * H],
                                   dqkv[:, 2 * H:], batch=B, num_heads=heads, nq=Q, nk=Q,
                                   dropout=site(li, K.KIND_SELF_PROBS))
            dh_in, g["w_qkv"], g["b_qkv"] = lin_bwd(dqkv, t["h_in"], L["w_qkv"])
            dy, dy2 = dpre1, dh_in
            layer_grads[li] = g
            ctx.tape[li] = None                           # free this layer's activations
            assert cursor[0] - bucket_lo == layer_numel(L)
            if on_ready is not None:
                on_ready(li, [v for v in g.values()], flat[bucket_lo:cursor[0]])
        # ---- cross-attention K/V projection of all layers (one GEMM in the forward pass)
        g_kv_w = g_kv_b = None
        kv_lo = cursor[0]
        if n_cross:
            _, g_kv_w, g_kv_b = lin_bwd(d_kv_all, ctx.enc2, pk["w_kv_all"], need_dx=False)
        # ---- query-token LayerNorm (batch-invariant): sum over the batch, then a [Q, H] LayerNorm backward (torch)
        dh0 = dy.float() + dy2.float()
        if drop is not None:
            dh0 = ops.dropout_backward(dh0.to(torch.bfloat16), (drop[0], drop[1], K.SITE_EMBEDDINGS) + tuple(drop[2:])).float()
        dh0 = dh0.view(B, Q, H).sum(dim=0)
        with torch.enable_grad():
            q0 = ctx.q0.clone().requires_grad_(True)
            eg = pk["emb_g"].clone().requires_grad_(True)
            eb = pk["emb_b"].clone().requires_grad_(True)
            torch.nn.functional.layer_norm(q0, (H,), eg, eb, eps).backward(dh0)
        # ---- scatter into the order of _live_params()
        grads = [eg.grad, eb.grad]
        for li, L in enumerate(pk["layers"]):
            g = layer_grads[li]
            wq, wk, wv = g["w_qkv"][:H], g["w_qkv"][H:2 * H], g["w_qkv"][2 * H:]
            bq, bk, bv = g["b_qkv"][:H], g["b_qkv"][H:2 * H], g["b_qkv"][2 * H:]
            grads += [wq, bq, wk, bk, wv, bv, g["w_o"], g["b_o"], g["ln1_g"], g["ln1_b"]]
            if L["cross"]:
                off = L["kv_slot"] * 2 * H
                grads += [g["w_qc"], g["b_qc"], g_kv_w[off:off + H], g_kv_b[off:off + H], g_kv_w[off + H:off + 2 * H],
                          g_kv_b[off + H:off + 2 * H], g["w_oc"], g["b_oc"], g["ln2_g"], g["ln2_b"]]
            grads += [g["w_1"], g["b_1"], g["w_2"], g["b_2"], g["ln3_g"], g["ln3_b"]]
        d_query = q0.grad.view(1, Q, H)
        if on_ready is not None:
            if n_cross:
                on_ready(-1, [g_kv_w, g_kv_b], flat[kv_lo:cursor[0]])
            on_ready(-3, [eg.grad, eb.grad, d_query])
        if on_finish is not None:
            on_finish()        # every bucket reduced and written back before autograd sees the gradients
        return (None, None, None, None, d_query) + tuple(grads)


class QFormerLoss(torch.nn.Module):
    """Drop-in for the reference's QFormerLoss (training/item_qformer_training.py:41-56): same constructor defaults
    (`reconstruction_weight=1.0, contrastive_weight=0.5, margin=0.5`), same call
    `criterion(model_output, input_embeddings, pos_rep, neg_rep, attention_mask)` with `model_output` / `input_embeddings`
    dicts (`reconstructed_fields`, `item_representation` / `field_embeddings`) and the same 3-tuple
    `(total, masked_recon_loss, cont_loss)`.  Masked MSE summed over fields and embedding dimension, divided by the number
    of valid FIELDS (:50-52 - not by fields x dimension); TripletMarginLoss(margin, p=2, eps=1e-6, mean) on
    item_representation (:45, :54).  Stays in PyTorch (0.003 % of the step's FLOPs); fp32."""

    def __init__(self, reconstruction_weight=1.0, contrastive_weight=0.5, margin=0.5):
        super().__init__()
        self.recon_w = reconstruction_weight
        self.cont_w = contrastive_weight
        self.margin = margin

    def forward(self, model_output, input_embeddings, pos_rep, neg_rep, attention_mask):
        rec = model_output["reconstructed_fields"].float()
        target = input_embeddings["field_embeddings"].float()
        mask = attention_mask.to(rec.dtype)
        masked_recon_loss = (((rec - target) ** 2) * mask.unsqueeze(-1)).sum() / mask.sum()
        cont_loss = torch.nn.functional.triplet_margin_loss(model_output["item_representation"].float(), pos_rep.float(),
                                                            neg_rep.float(), margin=self.margin, p=2)
        return self.recon_w * masked_recon_loss + self.cont_w * cont_loss, masked_recon_loss, cont_loss


def qformer_loss(outputs, field_embeddings, attention_mask, pos_rep=None, neg_rep=None, recon_weight: float = 1.0,
                 contrastive_weight: float = 0.5, margin: float = 0.5):
    """Total loss of `QFormerLoss` (reference defaults) as a plain function; without positive / negative representations
    only the reconstruction term (what a validation pass logs, training/item_qformer_training.py:150-155)."""
    if pos_rep is None or neg_rep is None:
        rec = outputs["reconstructed_fields"].float()
        mask = attention_mask.to(rec.dtype)
        return recon_weight * (((rec - field_embeddings.float()) ** 2) * mask.unsqueeze(-1)).sum() / mask.sum()
    return QFormerLoss(recon_weight, contrastive_weight, margin)(outputs, {"field_embeddings": field_embeddings},
                                                                pos_rep, neg_rep, attention_mask)[0]


class GradientAllReducer:
    """Data-parallel gradient averaging for the training step.  The backbone's backward
    calls `layer_ready` as soon as a layer's parameter gradients are complete; each call starts an asynchronous
    all-reduce of that layer's bucket (NCCL on GPUs, gloo in the CPU tests), so communication overlaps the backward of
    the remaining layers; `finish` (called at the end of the backbone's backward) waits and writes the results back in
    place.  Head gradients (2.1 M parameters) are reduced after backward with `reduce_params`.  The 132.5 M
    never-executed parameters (text branch, word / position embeddings) have no gradients and are never communicated.

    bucket_dtype: the WIRE dtype.  fp32: the backbone's flat fp32 gradient buffer is reduced in place (no copies).
    bf16: a bucket is cast to bf16 (one streaming pass), reduced, and cast back into the fp32 gradients - half the
    bytes over NVLink (357 MB instead of 714 MB per step); the accumulation inside a rank stays fp32, only the
    exchanged partial sums are rounded (2^-9 relative per rank).
    Averaging: NCCL reduces with ReduceOp.AVG (no separate divide pass); gloo has no AVG: SUM, then one divide.
    Every call is stream-ordered and allocates through torch's caching allocator only, so a step that contains these
    collectives can be captured into a CUDA graph (TrainStepGraph(reducer=...))."""

    def __init__(self, group=None, bucket_dtype: torch.dtype = torch.float32):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        ready = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if ready else 1
        self.bucket_dtype = bucket_dtype
        self.avg_in_collective = bool(ready and dist.get_backend(group) == "nccl")
        self.pending = []
        self.bytes_reduced = 0

    def attach(self, backbone):
        backbone.grad_ready_hook = self.layer_ready
        backbone.grad_finish_hook = self.finish
        return self

    def _start(self, wire: torch.Tensor, async_op: bool = True):
        op = self.dist.ReduceOp.AVG if self.avg_in_collective else self.dist.ReduceOp.SUM
        self.bytes_reduced += wire.numel() * wire.element_size()
        return self.dist.all_reduce(wire, op=op, group=self.group, async_op=async_op)

    def _settle(self, wire: torch.Tensor):
        if not self.avg_in_collective:
            wire.div_(self.world)

    def layer_ready(self, layer_index: int, tensors: List[torch.Tensor], flat: Optional[torch.Tensor] = None):
        """`flat`: the tensors are views that tile this contiguous buffer exactly (the backbone's gradient buffer) - with
        a matching wire dtype it is reduced in place, without the pack / unpack copies; with a bf16 wire it is cast once."""
        tensors = [t for t in tensors if t is not None]
        if self.world == 1 or not tensors:
            return
        if flat is not None and flat.dtype == self.bucket_dtype:
            self.pending.append((self._start(flat), flat, None, None))
        elif flat is not None:
            wire = flat.to(self.bucket_dtype)
            self.pending.append((self._start(wire), wire, flat, None))
        else:
            wire = torch.cat([t.reshape(-1).to(self.bucket_dtype) for t in tensors])
            self.pending.append((self._start(wire), wire, None, tensors))

    def finish(self):
        for work, wire, flat, tensors in self.pending:
            work.wait()
            self._settle(wire)
            if flat is not None:
                flat.copy_(wire)
            if tensors is not None:
                views, off = [], 0
                for t in tensors:
                    n = t.numel()
                    views.append(wire[off:off + n].view_as(t))
                    off += n
                torch._foreach_copy_(tensors, views)
        self.pending = []

    def reduce_params(self, params):
        """Synchronous averaging of .grad of parameters whose gradients are produced outside the backbone."""
        grads = [p.grad for p in params if p.grad is not None]
        self.layer_ready(-2, grads)
        self.finish()

    @staticmethod
    def alias_buckets(tensors: List[torch.Tensor], min_cover: float = 0.9):
        """Split `tensors` into (buckets, rest): a bucket is ONE flat tensor that aliases the storage range spanned by
        several of the tensors (contiguous views of one buffer that cover at least `min_cover` of that range, e.g. the
        backbone's flat gradient buffer behind autograd's `.grad` views) - reducing it in place reduces all of them
        without pack / unpack copies; `rest` are the tensors that stand alone."""
        groups = {}
        for t in tensors:
            groups.setdefault((t.untyped_storage().data_ptr(), t.dtype), []).append(t)
        buckets, rest = [], []
        for ts in groups.values():
            if len(ts) < 2 or any(not t.is_contiguous() for t in ts):
                rest += ts
                continue
            lo = min(t.storage_offset() for t in ts)
            hi = max(t.storage_offset() + t.numel() for t in ts)
            if sum(t.numel() for t in ts) < min_cover * (hi - lo):
                rest += ts
                continue
            buckets.append(ts[0].new_empty(0).set_(ts[0].untyped_storage(), lo, (hi - lo,)))
        return buckets, rest

    def reduce_tensors(self, tensors: List[torch.Tensor]):
        """Average `tensors` over the ranks in place.  Tensors that are views of one buffer are reduced as that buffer
        (one all-reduce; in place when the wire dtype matches, through one cast otherwise); the others are packed into
        one flat bucket, reduced and unpacked with a single multi-tensor copy.  Used behind a CUDA-graph step that was
        captured without collectives, whose gradients all exist when the graph has run."""
        tensors = [t for t in tensors if t is not None]
        if self.world == 1 or not tensors:
            return
        buckets, rest = self.alias_buckets(tensors)
        for bkt in buckets:
            self.layer_ready(-4, [bkt], bkt)
        if rest:
            self.layer_ready(-5, rest)
        self.finish()


class TrainStepGraph:
    """The reference's training step (training/item_qformer_training.py:117-131: forward, QFormerLoss, backward)
    captured ONCE into a CUDA graph and replayed per step.

    Why: one step is ~435 kernel launches and the Python host needs 35-38 ms to enqueue them - as long as the GPU needs
    for a 1024-item batch, and several times longer than the GPU needs for the 128-item per-GPU batch of an 8-GPU
    data-parallel run.  A replay is one launch.  What makes the step capturable:
      * every kernel of the C ABI is enqueued on torch's current stream and allocates nothing; the tensor maps are
        encoded on the host at capture time from addresses that stay valid (torch's graph-private memory pool);
      * dropout masks are a pure function of (seed, site, element); the seed frozen into the captured launch parameters
        is offset by a DEVICE-resident counter that the graph itself advances (`seed_offset`, include/unirec_b200.h),
        so every replay draws fresh masks and the backward pass of the same replay sees the same ones;
      * the bf16 weight packing of the backbone and the heads is part of the graph, so a replay always starts from
        the current fp32 master weights (the optimizer updates them in place between replays).
    Gradients land in static tensors (`.grads`); after `step()` they are attached to the parameters' `.grad`, so the
    caller averages them over the ranks (`GradientAllReducer.reduce_tensors(graph.grad_tensors())`) and runs the
    optimizer as usual.  Do not call `zero_grad(set_to_none=False)` between steps - a replay overwrites the gradients.
    Precondition (torch's, not ours): no autograd graph of an EARLIER eager step of the same model may still be alive
    when the graph is captured (e.g. a kept `loss` tensor) - it keeps the parameters' AccumulateGrad nodes bound to the
    eager stream and the capture is invalidated (`cudaErrorStreamCaptureInvalidated`); `del loss` first.
    """

    SEED_STRIDE = 1024     # > the number of forward passes in one step: replays never reuse an effective seed

    def __init__(self, model, field_embeddings: torch.Tensor, attention_mask: torch.Tensor, *,
                 faithful: bool = False, loss_kwargs: Optional[dict] = None, warmup: int = 3,
                 reducer: Optional["GradientAllReducer"] = None):
        """field_embeddings [B, F, E] / attention_mask [B, F]: example inputs (shape, dtype and device of every later
        step).  faithful: also run the two no-grad train-mode forwards that produce the positive / negative
        representations (item_qformer_training.py:122-125) inside the graph; otherwise they are inputs of `step`.
        reducer (data-parallel runs): the gradient averaging becomes PART OF THE GRAPH - the backbone's backward hands
        every layer's gradient bucket to `reducer.layer_ready` as soon as it is complete, the asynchronous NCCL all-reduce
        of layer l runs on NCCL's stream while the captured backward of layers l-1 .. 0 continues (the fork / join of the
        two streams is captured as graph dependencies), the head gradients follow at the end; after a replay every
        gradient is already averaged and the caller runs the optimizer directly.  Without it the caller reduces
        `grad_tensors()` behind the replay (one bucket, nothing to overlap with)."""
        if not field_embeddings.is_cuda:
            raise RuntimeError("TrainStepGraph: inputs must be CUDA tensors (unirec_b200 has no CPU path)")
        dev = field_embeddings.device
        self.model, self.faithful = model, faithful
        self.reducer = reducer if (reducer is not None and reducer.world > 1) else None
        self.loss_kwargs = dict(loss_kwargs or {})
        B, E = field_embeddings.shape[0], model.item_representation_head.out_features
        self.fields = field_embeddings.detach().clone()
        self.mask = attention_mask.detach().clone()
        if faithful:
            # the positive / negative items are masked with THEIR OWN attention masks (item_qformer_training.py:123-124)
            self.fields_pos = torch.zeros_like(self.fields)
            self.fields_neg = torch.zeros_like(self.fields)
            self.mask_pos = self.mask.clone()
            self.mask_neg = self.mask.clone()
        else:
            self.pos = torch.zeros(B, E, device=dev)
            self.neg = torch.zeros(B, E, device=dev)
        self.seed_offset = torch.zeros(1, dtype=torch.int64, device=dev)
        self.params = [p for p in model.parameters() if p.requires_grad]
        in_backbone = {id(p) for p in model.qformer.parameters()} | {id(model.query_embeddings)}
        self.head_params = [p for p in self.params if id(p) not in in_backbone]
        saved_offset = model.dropout_seed_offset
        saved_hooks = (getattr(model.qformer, "grad_ready_hook", None), getattr(model.qformer, "grad_finish_hook", None))
        model.dropout_seed_offset = self.seed_offset
        if self.reducer is not None:          # collectives inside the graph (the warm-up passes create the communicator)
            model.qformer.grad_ready_hook, model.qformer.grad_finish_hook = self.reducer.layer_ready, self.reducer.finish
        else:
            model.qformer.grad_ready_hook = model.qformer.grad_finish_hook = None
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 1)):       # lazy one-time work (function attributes, caches) happens here
                    self._clear_grads()
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            self._clear_grads()
            # the bf16 weight packs cached by the warm-up passes live outside the graph: drop them so that the packing
            # kernels are captured and every replay re-packs the CURRENT fp32 master weights
            model.qformer._pack = model.qformer._pack_key = None
            model._head_pack = model._head_key = None
            self.graph = torch.cuda.CUDAGraph()
            bytes_before = self.reducer.bytes_reduced if self.reducer is not None else 0
            # with collectives inside, NCCL's watchdog thread may query events while this thread captures: only calls
            # made by THIS thread may invalidate the capture
            mode = {"capture_error_mode": "thread_local"} if self.reducer is not None else {}
            with torch.cuda.graph(self.graph, **mode):
                self.seed_offset.add_(self.SEED_STRIDE)
                self.loss = self._body()
            # bytes every replay puts on the wire (the reducer's own counter only sees the capture)
            self.allreduce_bytes_per_step = (self.reducer.bytes_reduced - bytes_before) if self.reducer is not None else 0
        finally:
            model.dropout_seed_offset = saved_offset
            model.qformer.grad_ready_hook, model.qformer.grad_finish_hook = saved_hooks
        self.grads = [(p, p.grad) for p in self.params if p.grad is not None]
        self.replays = 0

    def _clear_grads(self):
        for p in self.params:
            p.grad = None

    def _body(self):
        model = self.model
        out = model(self.fields, self.mask)
        if self.faithful:
            with torch.no_grad():
                p_rep = model(self.fields_pos, self.mask_pos)["item_representation"]
                n_rep = model(self.fields_neg, self.mask_neg)["item_representation"]
        else:
            p_rep, n_rep = self.pos, self.neg
        loss = qformer_loss(out, self.fields, self.mask, p_rep, n_rep, **self.loss_kwargs)
        loss.backward()
        if self.reducer is not None:
            self.reducer.reduce_params(self.head_params)
        return loss.detach()

    def grad_tensors(self) -> List[torch.Tensor]:
        return [g for _, g in self.grads]

    def step(self, field_embeddings: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
             pos: Optional[torch.Tensor] = None, neg: Optional[torch.Tensor] = None,
             pos_mask: Optional[torch.Tensor] = None, neg_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Copies the batch into the graph's static inputs (device-to-device, on the current stream), replays the graph
        and returns the loss (a static tensor: read it before the next step).  `pos` / `neg` are the positive / negative
        REPRESENTATIONS [B, E] (faithful=False) or their FIELD EMBEDDINGS [B, F, E] (faithful=True); in the faithful step
        `pos_mask` / `neg_mask` [B, F] are those items' own attention masks (item_qformer_training.py:123-124) and are
        required whenever `pos` / `neg` are given."""
        if self.faithful and ((pos is not None and pos_mask is None) or (neg is not None and neg_mask is None)):
            raise ValueError("TrainStepGraph(faithful=True).step: pass pos_mask / neg_mask with pos / neg (the reference "
                             "masks the positive and negative items with their own attention masks)")
        if not self.faithful and (pos_mask is not None or neg_mask is not None):
            raise ValueError("TrainStepGraph(faithful=False).step: pos / neg are representations and take no mask")
        self.fields.copy_(field_embeddings, non_blocking=True)
        if attention_mask is not None:
            self.mask.copy_(attention_mask, non_blocking=True)
        if pos is not None:
            (self.fields_pos if self.faithful else self.pos).copy_(pos, non_blocking=True)
        if neg is not None:
            (self.fields_neg if self.faithful else self.neg).copy_(neg, non_blocking=True)
        if pos_mask is not None:
            self.mask_pos.copy_(pos_mask, non_blocking=True)
        if neg_mask is not None:
            self.mask_neg.copy_(neg_mask, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        for p, g in self.grads:
            p.grad = g
        return self.loss
