#!/usr/bin/env python
"""Benchmark of the nested Q-Former encode-and-rank hot path (BASELINE.json north_star / config 5).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" of the headline metric is one pass of the nested user path
over one batch of synthetic users: gather 50 history items x 32 query tokens from the resident item-token
table (+ sinusoidal PE) -> UserQFormer (64 queries x 1600 keys x 4 layers) -> prediction head -> pooled
scoring vector -> cosine scoring against the 1M-row candidate pool -> top-100.  The item-token table and
the candidate pool are produced beforehand by the item Q-Former over 1M synthetic items (config 3,
item-range sharded); that run is timed too and reported as the "items" block (items/sec).

Scaling is weak: every rank encodes --users-per-gpu users per step; user vectors are all-gathered, each
rank scores all users against its 1/N of the candidate rows, per-rank top-100 lists are all-gathered and
merged.  value = (N * users_per_gpu * K) / max-over-ranks device time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "users/sec (nested item->user Q-Former encode + top-100 rank over a 1M-item pool)"
FLOPS_PER_USER = 36_173_774_848          # SURVEY.md section 8d (user Q-Former + head, S = 1600)
FLOPS_PER_ITEM = 10_882_646_016          # SURVEY.md section 8d (item Q-Former -> query_outputs)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pool-items", type=int, default=1_000_000)
    ap.add_argument("--users-per-gpu", type=int, default=4096)
    ap.add_argument("--item-batch", type=int, default=4096)
    ap.add_argument("--history", type=int, default=50)
    ap.add_argument("--fused-gather", type=int, default=0, help="1: the user K/V projection gathers the history tokens from "
                    "the item-token table itself (no materialised user sequence; SURVEY 8f-2)")
    ap.add_argument("--fused-kv", type=int, default=0, help="1: the user cross-attention projects K / V inside the attention "
                    "kernel (csrc/kv_attention_fused.cu): no K/V buffer, one chunk of users per step")
    ap.add_argument("--kv-gb", type=float, default=0.0, help="user Q-Former: bytes of cross-attention K/V materialised per "
                    "chunk of users, in GiB (0 = the module's default)")
    ap.add_argument("--layer-major", type=int, default=0, help="1: user Q-Former in layer-major order (one encoder call per "
                    "step, K/V of ONE layer materialised per chunk of users inside it); 0: chunk-major (K/V of all 4 layers "
                    "per 512-user chunk, the sequence is read once)")
    ap.add_argument("--fold-ln", type=int, default=-1, help="1 / 0: LayerNorms of both encoders folded into the GEMMs around "
                    "them (no LayerNorm kernels between the GEMMs) / separate streaming LayerNorm kernels; -1: module default")
    ap.add_argument("--exchange", default="alltoall", choices=["alltoall", "allgather"], help="N > 1: how the per-rank top-k "
                    "lists travel (alltoall: every rank merges and returns its own users; allgather: every rank merges all)")
    ap.add_argument("--top-k", type=int, default=100)
    ap.add_argument("--cpu-users", type=int, default=32, help="users in the bounded CPU-baseline sample (timed); the "
                    "top-k parity check against the GPU result uses the first 8 of them")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernels-alone", action="store_true", help="skip the per-kernel burst timings (kernels_alone)")
    ap.add_argument("--train-batch", type=int, default=1024, help="GLOBAL batch of the cfg-2 training block (0 = skip)")
    ap.add_argument("--train-steps", type=int, default=4)
    ap.add_argument("--train-dropout", type=float, default=0.2, help="dropout of the cfg-2 training block "
                    "(reference default 0.2, models/qformer_utils.py:19)")
    ap.add_argument("--train-wire", default="bf16", choices=["bf16", "fp32"], help="dtype of the gradient buckets on the "
                    "wire (bf16: cast once per bucket, half the NVLink bytes; accumulation inside a rank stays fp32)")
    ap.add_argument("--train-only", action="store_true", help="run only the cfg-2 training block (diagnostics)")
    ap.add_argument("--profile-range", default="", choices=["", "items", "users", "train"],
                    help="bracket that timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return ap.parse_args()


def kernels_alone(dev, pk, pooled, k):
    """The HBM-bound kernels of the path timed ALONE (short bursts after an idle pause, CUDA events per launch, median of 9):
    the north star's per-kernel targets (attention and scoring >= 70 % of the HBM peak) next to the in-step figures of
    `roofline.other_kernels`, which are taken under the power cap of the 160 ms step (SM clock ~1.4 GHz instead of 1.9).
    Bytes are the algorithmic ones of SURVEY.md 8(d)."""
    import time
    import torch
    from unirec_b200 import ops
    H, heads = 1024, 16
    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(77)
    rnd = lambda *shape: (torch.randn(*shape, device=dev, generator=g) * 0.5).to(bf)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        time.sleep(0.3)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(9)]
        for a, b in ev:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return sorted(a.elapsed_time(b) for a, b in ev)[4]

    out = {}

    def hbm(name, ms, nbytes):
        gbs = nbytes / ms / 1e6
        out[name] = {"ms": ms, "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "bound": "hbm"}

    Bu, Q, S = 512, 64, 1600                                  # user cross-attention of one K/V chunk, one layer
    qc, kv, mask = rnd(Bu * Q, H), rnd(Bu * S, 2 * H), torch.ones(Bu, S, device=dev)
    hbm("attention_user_cross_512x64x1600",
        timed(lambda: ops.attention(qc, kv[:, :H], kv[:, H:], batch=Bu, num_heads=heads, nq=Q, nk=S, key_mask=mask)),
        (2 * Q + 2 * S) * Bu * H * 2)
    del kv
    qkv = rnd(2048 * 64, 3 * H)
    hbm("attention_user_self_2048x64x64",
        timed(lambda: ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], batch=2048, num_heads=heads, nq=64, nk=64)),
        4 * 2048 * 64 * H * 2)
    del qkv
    Bi = 8192
    qkv = rnd(Bi * 32, 3 * H)
    hbm("attention_item_self_8192x32x32",
        timed(lambda: ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], batch=Bi, num_heads=heads, nq=32, nk=32)),
        4 * Bi * 32 * H * 2)
    qi, kvi, mi = qkv[:, :H], rnd(Bi * 14, 2 * H), torch.ones(Bi, 14, device=dev)
    hbm("attention_item_cross_8192x32x14",
        timed(lambda: ops.attention(qi, kvi[:, :H], kvi[:, H:], batch=Bi, num_heads=heads, nq=32, nk=14, key_mask=mi)),
        (2 * 32 + 2 * 14) * Bi * H * 2)
    del qkv, kvi
    N = pooled.shape[0]
    cinv = ops.inv_l2_norm(pooled)
    u128 = rnd(128, H)
    hbm(f"score_topk_128_users_x_{N}", timed(lambda: ops.score_topk(u128, pooled, k, cand_inv=cinv)), N * H * 2)
    u4k = rnd(4096, H)
    ms = timed(lambda: ops.score_topk(u4k, pooled, k, cand_inv=cinv))
    tf = 2.0 * 4096 * N * H / ms / 1e9
    out[f"score_topk_4096_users_x_{N}"] = {"ms": ms, "achieved": tf, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                                             "frac": tf / pk["bf16_burst"], "bound": "tensor (burst peak: kernel timed alone)"}
    return out


def config_dict(args, n_gpus):
    return {
        "workload": "cfg5 nested item->user Q-Former encode + cosine top-100 over a 1M-item candidate pool "
                    "(item-token table + pool produced by cfg3 item-token generation)",
        "users_per_gpu_per_step": args.users_per_gpu, "global_users_per_step": args.users_per_gpu * n_gpus,
        "user_chunk_kv_gib": args.kv_gb if args.kv_gb > 0 else "module default",
        "user_sequence": "gathered inside the K/V projection" if args.fused_gather else "materialised per chunk",
        "user_cross_attention": ("K/V projected inside the attention kernel (no K/V in HBM)" if args.fused_kv else
                                 "layer-major: K/V of one layer materialised per chunk of users (<= 14 GiB), then attention"
                                 if args.layer_major else "K/V of all layers materialised per chunk, then attention"),
        "layernorm": "module default" if args.fold_ln < 0 else ("folded into the GEMMs" if args.fold_ln else "streaming kernels"),
        "history_items": args.history, "tokens_per_item": 32, "keys_per_user": args.history * 32,
        "user_qformer": "4 layers x 64 queries, hidden 1024, 16 heads, FFN 4096, cross-attn every layer",
        "item_qformer": "12 layers x 32 queries, 14 fields x 1024, cross-attn every 2nd layer",
        "pool_items": args.pool_items, "top_k": args.top_k, "item_batch": args.item_batch,
        "parallelism": f"users dp{n_gpus}; candidate rows sharded /{n_gpus}; all-gather of user vectors, "
                       + ("all-to-all of the per-rank top-k lists (int32 index + shard base), every rank merges its own users"
                          if args.exchange == "alltoall" else "all-gather + merge of top-k on every rank"),
        "l2": "inputs larger than L2 every step (user sequences 13.4 GB, candidate pool 2 GB / n_gpus, "
              "item field batches generated on device per chunk: 1 M distinct items)",
    }


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# a-priori tolerances: cosine scores of unit vectors, D = 1024 - bf16 encoders vs the fp32 oracle / the
# same bf16 inputs scored in fp32 by another implementation
SCORE_TOL_BF16 = 4e-3
SCORE_TOL_SAME_INPUTS = 1e-4


def parity_rows(n_users: int, want: int):
    """Rows of a timed call to check: the chunk boundaries of the 512-user chunks, the middle, the last user, then an
    even spread."""
    fixed = [r for r in (0, 511, 512, 513, 1023, 1024, 2047, 2048, n_users - 1) if 0 <= r < n_users]
    rows = list(dict.fromkeys(fixed))[:want]
    step = max(n_users // max(want, 1), 1)
    for r in range(step // 2, n_users, step):
        if len(rows) >= want:
            break
        if r not in rows:
            rows.append(r)
    return sorted(rows)


def topk_parity(got_s, got_i, ref_s, ref_i, full_scores, score_tol: float, tie_tol: float):
    """North-star parity of top-k lists: scores within `score_tol`; indices identical except where scores tie - a pick
    that is not in the reference list must score (in the reference's own full score row) within `tie_tol` of the
    reference's k-th score.  All CPU tensors; full_scores [B, N]."""
    B, k = got_i.shape
    overlap = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(got_i, ref_i)) / (B * k)
    # compare the scores of the picks themselves (the sorted lists pair up different items where ranks swap)
    picked = full_scores.gather(1, got_i.long())
    score_diff = float((got_s - picked).abs().max())
    outside = 0
    for u in range(B):
        kth = float(ref_s[u, -1])
        extra = set(got_i[u].tolist()) - set(ref_i[u].tolist())
        outside += sum(1 for j in extra if float(full_scores[u, j]) < kth - tie_tol)
    return {"users_checked": B, "score_max_abs_diff": score_diff, "score_tolerance": score_tol,
            "topk_overlap": overlap, "tie_tolerance": tie_tol, "picks_outside_tie_tolerance": outside,
            "ok": bool(score_diff <= score_tol and outside == 0)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_user_rank_sample(usd, tokens_cpu, history, lengths, cands_cpu, k, heads=16, predict=32):
    """One pass of the oracle (CPU restatement of the reference) over a bounded sample of users."""
    import torch
    from oracle import qformer_oracle as O
    with torch.no_grad():
        seq, mask = O.build_user_sequences(tokens_cpu, history, lengths)
        pred = O.user_qformer_forward(usd, seq, mask, num_heads=heads, num_item_tokens_to_predict=predict)
        u = O.pooled_scoring_vector(pred)
        return O.cosine_topk(u, cands_cpu, k)


def torch_eager_gpu_sample(item, user, fpool, tokens, pooled, hist, lengths, k, dev):
    """Context, not a baseline the contract asks for: the reference's own eager PyTorch code ON THE SAME B200 - the
    UNMODIFIED reference modules (oracle/_ref) moved to the GPU, or the oracle's functional restatement when that copy is
    absent - in fp32 exactly like the reference's default and under bf16 autocast.  Bounded samples."""
    import torch
    from oracle import qformer_oracle as O
    out = {}
    try:
        with torch.no_grad():
            isd = {k_: v.detach().float() for k_, v in item.state_dict().items()}
            usd = {k_: v.detach().float() for k_, v in user.state_dict().items()}
            xi = fpool[0, :256].float()
            mi = torch.ones(256, 14, device=dev, dtype=torch.long)
            Bu = 32
            tok_u = tokens[hist[:Bu].reshape(-1)].float().view(Bu, hist.shape[1], 32, -1)
            lens = lengths[:Bu].long()
            cands = pooled.float()
            ref = reference_models(isd, usd, dev)

            def timed(fn, n=2):
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / n

            if ref is not None:
                RR, r_item, r_user, r_pe = ref
                out["code"] = "UNMODIFIED reference modules (oracle/_ref) on the GPU"
                item_path = lambda: RR.item_tokens(r_item, xi, mi)
                user_path = lambda: RR.user_rank(r_user, r_pe, tok_u, lens, cands, k)
            else:
                out["code"] = "oracle restatement on the GPU (oracle/_ref not present)"
                hist_l = torch.arange(Bu * hist.shape[1], device=dev).view(Bu, -1)
                flat = tok_u.reshape(-1, 32, tok_u.shape[-1])
                with torch.device(dev):
                    seq, mask = O.build_user_sequences(flat, hist_l, lens)
                item_path = lambda: O.item_qformer_forward(isd, xi, None)

                def user_path():
                    with torch.device(dev):
                        pred = O.user_qformer_forward(usd, seq, mask, num_heads=16, num_item_tokens_to_predict=32)
                        return O.cosine_topk(O.pooled_scoring_vector(pred), cands, k)

            for name, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
                if ctx is not None:
                    ctx.__enter__()
                try:
                    ms_i = timed(item_path)
                    ms_u = timed(user_path)
                finally:
                    if ctx is not None:
                        ctx.__exit__(None, None, None)
                out[name] = {"items_per_s": 256 / (ms_i * 1e-3), "users_per_s": Bu / (ms_u * 1e-3)}
            out["sample"] = "256 items; 32 users (S=1600) + cosine top-k over the full pool; eager torch ops, no compile"
    except Exception as e:  # context only: never fail the bench on it
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    return out


def reference_models(item_sd, user_sd, device):
    """The UNMODIFIED reference modules (oracle/_ref or /root/reference through the shim), or None if not present."""
    try:
        from oracle import reference_runner as RR
        if not RR.available():
            return None
        item, user, pe = RR.load_models(item_sd, user_sd, device=device)
        return RR, item, user, pe
    except Exception as e:      # a broken copy must not take the bench down: fall back to the oracle port and say so
        sys.stderr.write(f"bench: reference modules unavailable ({type(e).__name__}: {e}); using the oracle port\n")
        return None


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores - the UNMODIFIED
    reference modules from oracle/_ref (kind "reference"; oracle/build_ref.py copies them verbatim at build time), or the
    oracle port if that copy is missing (kind "port").  All host threads, bounded sample per step."""
    if rank != 0:
        return
    import torch
    from unirec_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    B, H = args.cpu_users, args.history
    usd = synth.user_qformer_state_dict(seed=0, live_only=False)
    g = torch.Generator().manual_seed(0)
    n_tok = B * H
    tokens = torch.randn(n_tok, 32, 1024, generator=g)
    history = torch.arange(n_tok).view(B, H)
    lengths = torch.full((B,), H, dtype=torch.long)
    cands = torch.randn(args.pool_items, 1024, generator=g)
    ref = reference_models(None, usd, "cpu")
    if ref is not None:
        RR, _, user, pe = ref
        kind = "reference"

        def one_pass(tok, hist, lens, cnd):
            return RR.user_rank(user, pe, tok[hist], lens, cnd, args.top_k)
    else:
        kind = "port"

        def one_pass(tok, hist, lens, cnd):
            return cpu_user_rank_sample(usd, tok, hist, lens, cnd, args.top_k)
    for _ in range(max(1, min(args.warmup, 1))):
        one_pass(tokens[:H], history[:1], lengths[:1], cands[:4096])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass(tokens, history, lengths, cands)
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    sample = (f"{B} users/step x {args.steps} steps: sequence build + user Q-Former (S={H * 32}) + cosine top-"
              f"{args.top_k} over {args.pool_items} candidates, fp32, torch CPU, "
              + ("UNMODIFIED reference modules (oracle/_ref)" if kind == "reference" else "oracle port"))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "users/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "users/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ------------------------------------------------------------------------------- cfg 2: training block
def run_train_block(args, rank, world, dev, barrier, max_over_ranks, pk):
    """BASELINE config 2: item Q-Former training step (fwd + QFormerLoss + bwd + gradient all-reduce + AdamW),
    global batch args.train_batch split over the ranks (strong scaling), bf16 activations / fp32 master weights,
    train-mode dropout args.train_dropout (Philox masks inside the kernels).  Primary step = anchor forward + backward
    (positive / negative representations given); the "reference_faithful" step adds the two no-grad train-mode
    forwards of training/item_qformer_training.py:122-125.  Returns the "train" block of the JSON line."""
    import torch
    from unirec_b200 import _lib, ops
    from unirec_b200.modules import QFormerForItemRepresentation
    from unirec_b200.training import GradientAllReducer, qformer_loss
    Bg = args.train_batch
    Bl = Bg // world
    torch.manual_seed(0)
    with torch.device(dev):
        model = QFormerForItemRepresentation(num_fields=14, dropout=args.train_dropout).train()
    model.dropout_seed = 1000 + rank * 1_000_003
    opt = torch.optim.AdamW([p for p in model.parameters()], lr=1e-4, fused=True)
    wire = torch.bfloat16 if args.train_wire == "bf16" else torch.float32
    red = GradientAllReducer(bucket_dtype=wire).attach(model.qformer)
    head_params = (list(model.item_representation_head.parameters()) + list(model.reconstruction_head.parameters()) +
                   list(model.field_projection.parameters()))
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    nb = 3
    fields = torch.randn(nb, Bl, 14, 1024, device=dev, generator=gen)
    fields[:, :, 7, 768:] = 0
    mask = torch.ones(Bl, 14, device=dev)
    pos = torch.randn(Bl, 1024, device=dev, generator=gen)
    neg = torch.randn(Bl, 1024, device=dev, generator=gen)

    def step(i, faithful=False):
        out = model(fields[i % nb], mask)
        p_rep, n_rep = pos, neg
        if faithful:
            with torch.no_grad():
                p_rep = model(fields[(i + 1) % nb], mask)["item_representation"]
                n_rep = model(fields[(i + 2) % nb], mask)["item_representation"]
        loss = qformer_loss(out, fields[i % nb], mask, p_rep, n_rep)
        loss.backward()
        red.reduce_params(head_params)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for i in range(2):
        step(i)
    barrier()
    bytes0 = red.bytes_reduced
    step(2)
    eager_allreduce_bytes = red.bytes_reduced - bytes0          # one eager step's wire bytes
    l0 = _lib.launch_count()
    if world == 1:
        ops.start_timing()      # per-launch CUDA events cost the host ~5 ms per step: only for the 1-GPU roofline
    if args.profile_range == "train":
        torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prof = None
    if os.environ.get("BENCH_PROFILE_TRAIN"):          # diagnostics: host-side profile of the steady-state loop
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    t_host0 = time.perf_counter()
    for i in range(args.train_steps):
        loss = step(i)
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.train_steps
    if prof is not None:
        import io
        import pstats
        prof.disable()
        buf = io.StringIO()
        pstats.Stats(prof, stream=buf).sort_stats("tottime").print_stats(40)
        pstats.Stats(prof, stream=buf).sort_stats("cumulative").print_stats(40)
        open(os.environ["BENCH_PROFILE_TRAIN"], "w").write(buf.getvalue())
    e1.record()
    barrier()
    if args.profile_range == "train":
        torch.cuda.profiler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.train_steps
    stats = ops.stop_timing() if world == 1 else {}
    launches = (_lib.launch_count() - l0) // args.train_steps
    step(0, faithful=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.train_steps):
        step(i, faithful=True)
    f1.record()
    barrier()
    ms_faithful = max_over_ranks(f0.elapsed_time(f1)) / args.train_steps
    gemm = stats.get("gemm")
    ach = gemm[1] / (gemm[2] * 1e-3) / 1e12 if gemm and gemm[2] > 0 else None
    eager_ms, eager_faithful_ms, eager_host_ms = ms, ms_faithful, host_ms
    final_loss = float(loss.detach())
    # an autograd graph kept alive by `loss` keeps the parameters' AccumulateGrad nodes - bound to the stream of the
    # eager steps - alive, and capture on another stream would be invalidated by them
    del loss
    import gc
    gc.collect()

    # ---- the same step replayed from a CUDA graph (training.TrainStepGraph): one launch per step instead of ~435, the
    # gradient all-reduce (one flat bucket) and the fused AdamW stay outside the graph
    graph_info = {"used": False}
    try:
        from unirec_b200.training import TrainStepGraph
        opt.zero_grad(set_to_none=True)

        def timed_graph(faithful, captured_collectives):
            # data-parallel runs: the per-layer gradient all-reduces are captured INTO the graph (they overlap the rest
            # of the captured backward on NCCL's stream); otherwise one flat bucket is reduced behind the replay
            tg = TrainStepGraph(model, fields[0], mask, faithful=faithful,
                                reducer=red if captured_collectives else None)

            def gstep(i):
                if faithful:
                    loss_ = tg.step(fields[i % nb], mask, fields[(i + 1) % nb], fields[(i + 2) % nb], mask, mask)
                else:
                    loss_ = tg.step(fields[i % nb], mask, pos, neg)
                if not captured_collectives:
                    red.reduce_tensors(tg.grad_tensors())
                opt.step()
                return loss_

            for i in range(2):
                gstep(i)
            barrier()
            b0 = red.bytes_reduced
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            t_h = time.perf_counter()
            for i in range(args.train_steps):
                loss_ = gstep(i)
            h_ms = (time.perf_counter() - t_h) * 1e3 / args.train_steps
            g1.record()
            barrier()
            ms_ = max_over_ranks(g0.elapsed_time(g1)) / args.train_steps
            val = float(loss_)
            wire_bytes = tg.allreduce_bytes_per_step if captured_collectives else \
                (red.bytes_reduced - b0) // max(args.train_steps, 1)
            opt.zero_grad(set_to_none=True)
            del tg
            torch.cuda.empty_cache()
            return ms_, h_ms, val, wire_bytes

        variants = {}
        if world > 1:
            try:
                variants["collectives_in_graph"] = timed_graph(False, True)
            except Exception as e:      # e.g. a NCCL / torch build that cannot capture collectives
                variants["collectives_in_graph_error"] = f"{type(e).__name__}: {e}"[:300]
                torch.cuda.synchronize()
        variants["reduce_after_replay"] = timed_graph(False, False)
        timed = {k: v for k, v in variants.items() if isinstance(v, tuple)}
        best = min(timed, key=lambda k: timed[k][0])
        g_ms, g_host_ms, g_loss, g_wire = timed[best]
        gf_ms, _, _, _ = timed_graph(True, best == "collectives_in_graph")
        graph_info = {"used": True, "ms_per_step": g_ms, "host_enqueue_ms_per_step": g_host_ms,
                      "faithful_ms_per_step": gf_ms, "final_loss": g_loss, "gradient_exchange": best,
                      "allreduce_bytes_per_step": g_wire,
                      "variants_ms_per_step": {k: (v[0] if isinstance(v, tuple) else v) for k, v in variants.items()},
                      "what": "forward + loss + backward replayed from one CUDA graph (device-resident dropout seed "
                              "offset); data-parallel: per-layer gradient buckets all-reduced INSIDE the graph on NCCL's "
                              "stream, overlapping the captured backward (or one flat bucket behind the replay, whichever "
                              "is faster); fused AdamW outside the graph"}
        if g_ms < eager_ms:       # the headline of the block is the faster mode; both are reported
            ms, ms_faithful, host_ms = g_ms, gf_ms, g_host_ms
        else:
            graph_info["note"] = ("eager is faster at this size: its per-layer all-reduce overlaps the backward pass and "
                                  "the host still keeps ahead of the GPU")
    except Exception as e:   # keep the eager numbers if capture is not possible on this box
        graph_info = {"used": False, "error": f"{type(e).__name__}: {e}"[:300]}
    items_per_sec = Bg / (ms * 1e-3)
    del model, opt
    torch.cuda.empty_cache()
    return {
        "metric": "items/sec (item Q-Former training step: fwd + QFormerLoss + bwd + grad all-reduce + AdamW)",
        "value": items_per_sec, "unit": "items/s", "global_batch": Bg, "per_gpu_batch": Bl, "ms_per_step": ms,
        "scaling": "strong", "dtype": "bf16 activations, fp32 master weights / gradients",
        "dropout": args.train_dropout, "host_enqueue_ms_per_step": host_ms,
        "reference_faithful_step": {"what": "anchor fwd+bwd + two no-grad train-mode forwards (positive, negative)",
                                    "ms_per_step": ms_faithful, "items_per_s": Bg / (ms_faithful * 1e-3)},
        "final_loss": final_loss, "gpu_launches_per_step": launches,
        "mode": "cuda_graph" if (graph_info.get("used") and ms < eager_ms) else "eager",
        "cuda_graph": graph_info,
        "eager": {"ms_per_step": eager_ms, "items_per_s": Bg / (eager_ms * 1e-3), "host_enqueue_ms_per_step": eager_host_ms,
                  "faithful_ms_per_step": eager_faithful_ms},
        "allreduce_bytes_per_step": (graph_info.get("allreduce_bytes_per_step", 0)
                                     if (graph_info.get("used") and ms < eager_ms) else eager_allreduce_bytes),
        "allreduce_wire_dtype": args.train_wire,
        "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": (ach / pk["bf16_sustained"]) if ach else None,
                     "share_of_step": (gemm[2] / (eager_ms * args.train_steps)) if gemm else None,
                     "timed_in": "eager steps (per-launch CUDA events cannot be recorded inside a graph replay)",
                     "end_to_end_frac_of_tensor_peak":
                         items_per_sec / world * 3 * FLOPS_PER_ITEM / (pk["bf16_sustained"] * 1e12)},
    }


# ------------------------------------------------------------------------------------------------- ours
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from unirec_b200 import _lib, ops
    from unirec_b200.modules import QFormerForItemRepresentation, UserQFormer
    from unirec_b200.pipeline import NestedRanker, shard_range

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU: unirec_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; stdout carries exactly one JSON line,
        # so route fd 1 to stderr while the communicator is created
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pk = peaks()
    if args.train_only:
        tb = run_train_block(args, rank, world, dev, barrier, max_over_ranks, pk)
        if rank == 0:
            print(json.dumps({"train": tb}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    torch.manual_seed(0)   # same random-init weights on every rank
    with torch.device(dev):
        item = QFormerForItemRepresentation(num_fields=14).eval()
        user = UserQFormer().eval()
    item.prelayernorm_dtype = torch.bfloat16
    user.prelayernorm_dtype = torch.bfloat16
    if args.kv_gb > 0:
        user.max_kv_bytes = int(args.kv_gb * (1 << 30))
    user.fused_kv_attention = bool(args.fused_kv)
    user.layer_major = bool(args.layer_major)
    if args.fold_ln >= 0:
        item.qformer.fold_layernorm = user.qformer.fold_layernorm = bool(args.fold_ln)

    # ------------------------------------------------------------------ stage A: item-token generation (cfg 3)
    N, Bi = args.pool_items, args.item_batch
    lo, hi = shard_range(N, rank, world)
    n_local = hi - lo
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool_chunks = 4
    fpool = torch.randn(pool_chunks, Bi, 14, 1024, device=dev, generator=gen)
    fpool[:, :, 7, 768:] = 0                     # CLIP ViT-L/14 768-d field zero-padded to 1024
    fmask = torch.ones(Bi, 14, device=dev, dtype=torch.long)
    tokens = torch.empty(N if world > 1 else n_local, 32, 1024, device=dev, dtype=torch.bfloat16)
    tok_local = tokens[lo:hi] if world > 1 else tokens
    pooled = torch.empty(n_local, 1024, device=dev, dtype=torch.bfloat16)
    chunks = [(c0, min(c0 + Bi, n_local)) for c0 in range(0, n_local, Bi)]

    def encode_chunk(ci):
        c0, c1 = chunks[ci]
        # every chunk gets its own field embeddings (SURVEY.md 8d cfg 3: generated on device per chunk): one in-place
        # normal_ over a rotating 235 MB buffer inside the timed loop (~0.3 % of a chunk's time).  Cycling over four fixed
        # batches instead made every item occur ~61 times in the 1 M pool - 61-fold ties at every rank of every top-100
        fbuf = fpool[ci % pool_chunks]
        fbuf.normal_(generator=gen)
        fbuf[:, 7, 768:] = 0                     # CLIP ViT-L/14 768-d field zero-padded to 1024
        tok = item.encode_query_tokens(fbuf[:c1 - c0], fmask[:c1 - c0], out_dtype=torch.bfloat16)
        tok_local[c0:c1].copy_(tok)
        pooled[c0:c1].copy_(ops.mean_tokens(tok))

    w_items = min(args.warmup, max(len(chunks) - 1, 0))
    for ci in range(w_items):
        encode_chunk(ci)
    barrier()
    l0 = _lib.launch_count()
    ops.start_timing()
    if args.profile_range == "items":
        torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for ci in range(w_items, len(chunks)):
        encode_chunk(ci)
    e1.record()
    barrier()
    if args.profile_range == "items":
        torch.cuda.profiler.stop()
    item_ms = max_over_ranks(e0.elapsed_time(e1))
    item_stats = ops.stop_timing()
    items_timed_local = sum(c1 - c0 for c0, c1 in chunks[w_items:])
    items_timed = items_timed_local * world if world > 1 else items_timed_local
    item_launches = _lib.launch_count() - l0
    items_per_sec = items_timed / (item_ms * 1e-3) if item_ms > 0 else 0.0

    if world > 1:   # replicate the token table (history gathers hit arbitrary items); setup, untimed
        for r in range(world):
            rlo, rhi = shard_range(N, r, world)
            step_rows = 8192
            for s0 in range(rlo, rhi, step_rows):
                dist.broadcast(tokens[s0:min(s0 + step_rows, rhi)], src=r)

    # ------------------------------------------------------------------ stage B: nested user encode + rank (cfg 5)
    Bu, Hh, k = args.users_per_gpu, args.history, args.top_k
    ranker = NestedRanker(user, tokens, pooled, k=k, index_base=lo, fused_gather=bool(args.fused_gather))
    hgen = torch.Generator(device=dev).manual_seed(99 + rank)
    hist_batches = [torch.randint(0, N, (Bu, Hh), device=dev, generator=hgen) for _ in range(3)]
    lengths = torch.full((Bu,), Hh, device=dev, dtype=torch.int32)

    local_res = args.exchange == "alltoall"
    for s in range(args.warmup):
        ranker(hist_batches[s % 3], lengths, local_result=local_res)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ops.start_timing()
    if args.profile_range == "users":
        torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        scores, idx = ranker(hist_batches[s % 3], lengths, local_result=local_res)
    e1.record()
    barrier()
    if args.profile_range == "users":
        torch.cuda.profiler.stop()
    # parity is checked on ROWS OF THE LAST TIMED CALL (chunk boundaries and the last user included), not on a side call
    timed_hist = hist_batches[(args.steps - 1) % 3]
    timed_scores, timed_idx = scores[:Bu].clone(), idx[:Bu].clone()      # rank 0's users are rows [0, Bu) of the result
    timed_u = ranker.last_user_vectors[:Bu].clone()                      # ... and of the gathered user vectors
    user_ms = max_over_ranks(e0.elapsed_time(e1))
    user_stats = ops.stop_timing()
    clocks = sampler.stop() if rank == 0 else None
    user_launches = _lib.launch_count() - l0
    users_per_sec = Bu * world * args.steps / (user_ms * 1e-3)

    # ------------------------------------------------------------------ e2e: host buffers in, host results out
    h_hist = [b.cpu().pin_memory() for b in hist_batches]
    h_len = lengths.cpu().pin_memory()
    out_rows = Bu if (local_res or world == 1) else Bu * world          # result rows a rank hands back to its host
    h_scores = torch.empty(out_rows, k, dtype=torch.float32).pin_memory()
    h_idx = torch.empty(out_rows, k, dtype=torch.int64).pin_memory()

    def e2e_step(s):
        d_hist = h_hist[s % 3].to(dev, non_blocking=True)
        d_len = h_len.to(dev, non_blocking=True)
        sc, ix = ranker(d_hist, d_len, local_result=local_res)
        h_scores.copy_(sc, non_blocking=True)
        h_idx.copy_(ix, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        e2e_step(s)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_users = Bu * world * args.steps / e2e_s
    h2d = (Bu * Hh * 8 + Bu * 4) * world        # whole job: every rank copies its users' histories in ...
    d2h = out_rows * k * 12 * world             # ... and its result rows (fp32 score + int64 index) out

    # items e2e: fp32 field embeddings from pinned host memory, bf16 tokens back to pinned host memory
    h_fields = fpool[0].cpu().pin_memory()
    h_fmask = fmask.cpu().pin_memory()
    h_tok = torch.empty(Bi, 32, 1024, dtype=torch.bfloat16).pin_memory()

    def item_e2e_step():
        x = h_fields.to(dev, non_blocking=True)
        m = h_fmask.to(dev, non_blocking=True)
        h_tok.copy_(item.encode_query_tokens(x, m, out_dtype=torch.bfloat16), non_blocking=True)
        torch.cuda.synchronize()

    item_e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e_items = 3
    for _ in range(n_e2e_items):
        item_e2e_step()
    barrier()
    item_e2e_serial_s = max_over_ranks(time.perf_counter() - t0)
    # the same host-to-host work through the public streamed loop: transfers of neighbouring chunks overlap the encode
    from unirec_b200.pipeline import generate_item_tokens_streamed
    n_e2e_items = 6
    h_fields_n = torch.empty(n_e2e_items * Bi, 14, 1024, dtype=torch.float32).pin_memory()
    for j in range(n_e2e_items):
        h_fields_n[j * Bi:(j + 1) * Bi].copy_(h_fields)
    h_fmask_n = h_fmask.repeat(n_e2e_items, 1).pin_memory()
    h_tok_n = torch.empty(n_e2e_items * Bi, 32, 1024, dtype=torch.bfloat16).pin_memory()
    generate_item_tokens_streamed(item, h_fields_n[:2 * Bi], h_fmask_n[:2 * Bi], h_tok_n[:2 * Bi], batch_size=Bi)
    barrier()
    t0 = time.perf_counter()
    generate_item_tokens_streamed(item, h_fields_n, h_fmask_n, h_tok_n, batch_size=Bi)
    barrier()
    item_e2e_s = max_over_ranks(time.perf_counter() - t0)
    del h_fields_n, h_tok_n

    # ------------------------------------------------------------------ parity at N > 1: merged lists vs brute force
    multi_parity = None
    if world > 1:
        # every rank contributes its candidate rows; rank 0 scores 16 of ITS users of the timed call against the whole
        # pool with plain torch ops (checker, untimed) and compares with the merged top-k the timed call returned
        from unirec_b200.pipeline import gather_rows
        full_pool = gather_rows(pooled)                                   # [N, 1024] bf16, rows in rank order = global ids
        if rank == 0:
            rows = parity_rows(Bu, 16)
            u = timed_u[rows].float()
            sims = torch.nn.functional.normalize(u, dim=-1) @ torch.nn.functional.normalize(full_pool.float(), dim=-1).t()
            ref_s, ref_i = torch.topk(sims, k, dim=-1)
            got_s, got_i = timed_scores[rows], timed_idx[rows]
            multi_parity = topk_parity(got_s.cpu(), got_i.cpu(), ref_s.cpu(), ref_i.cpu(), sims.cpu(),
                                       score_tol=SCORE_TOL_SAME_INPUTS, tie_tol=2 * SCORE_TOL_SAME_INPUTS)
            multi_parity["what"] = (f"rank 0: {len(rows)} users of the timed call (rows {rows[:6]}...) brute-forced with torch "
                                    f"fp32 against the all-gathered {full_pool.shape[0]}-row pool vs the merged top-{k} lists")
            del sims
        del full_pool

    # ------------------------------------------------------------------ CPU baseline inputs (rank 0, N == 1 only)
    # (the timed CPU passes run AFTER the training block: the oracle's 16 OpenMP threads keep spinning for a while and
    #  would slow the host thread that enqueues the training step)
    cpu_in = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rows = parity_rows(Bu, args.cpu_users)
        hist_c = timed_hist[rows]
        cpu_in = {"usd": {k_: v.detach().float().cpu() for k_, v in user.state_dict().items()},
                  "tok": tokens[hist_c.reshape(-1)].float().cpu(), "cands": pooled.float().cpu(),
                  "got_s": timed_scores[rows].cpu(), "got_i": timed_idx[rows].cpu(), "rows": rows,
                  "isd": {k_: v.detach().float().cpu() for k_, v in item.state_dict().items()},
                  "xi": fpool[0, :32].cpu()}

    # ------------------------------------------------------------------ eager PyTorch on the same GPU (context only)
    eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        eager = torch_eager_gpu_sample(item, user, fpool, tokens, pooled, hist_batches[0], lengths, k, dev)

    # ------------------------------------------------------------------ HBM-bound kernels timed alone (N == 1)
    alone = None
    if world == 1 and not args.no_kernels_alone:
        try:
            alone = kernels_alone(dev, pk, pooled, k)
        except Exception as e:                      # diagnostics only: never fail the bench line
            alone = {"error": repr(e)}

    # ------------------------------------------------------------------ cfg 2: training step
    train_block = None
    if args.train_batch > 0 and args.train_batch % world == 0:
        del ranker, tokens, tok_local, pooled, fpool
        torch.cuda.empty_cache()
        train_block = run_train_block(args, rank, world, dev, barrier, max_over_ranks, pk)

    # ------------------------------------------------------------------ CPU baseline (rank 0, N == 1 only)
    cpu_baseline, items_cpu = None, None
    if cpu_in is not None:
        from oracle import qformer_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        cores = torch.get_num_threads()
        rows = cpu_in["rows"]
        Bc = len(rows)
        usd, tok_c, cands_c = cpu_in["usd"], cpu_in["tok"], cpu_in["cands"]
        hist_local = torch.arange(Bc * Hh).view(Bc, Hh)
        len_c = torch.full((Bc,), Hh, dtype=torch.long)
        # the timed CPU arm: the UNMODIFIED reference modules when oracle/_ref is present, else the oracle port
        ref = reference_models(cpu_in["isd"], usd, "cpu")
        if ref is not None:
            RR, r_item, r_user, r_pe = ref
            kind = "reference"
            cpu_pass = lambda n: RR.user_rank(r_user, r_pe, tok_c[:n * Hh].view(n, Hh, 32, -1), len_c[:n], cands_c, k)
        else:
            kind = "port"
            cpu_pass = lambda n: cpu_user_rank_sample(usd, tok_c[:n * Hh], hist_local[:n], len_c[:n], cands_c, k)
        cpu_pass(1)                                                                            # warm-up
        t0 = time.perf_counter()
        ref_s, ref_i = cpu_pass(Bc)
        dt = time.perf_counter() - t0
        # parity of the TIMED GPU call against the fp32 CPU result, with tolerances fixed a priori: cosine
        # scores of unit vectors from bf16 encoders agree within SCORE_TOL_BF16; a GPU pick outside the oracle's list must
        # tie with the oracle's k-th score within 2 x that bound.  Checked on every sampled user (full score rows).
        full = O.cosine_scores(O.pooled_scoring_vector(
            O.user_qformer_forward(usd, *O.build_user_sequences(tok_c, hist_local, len_c), num_heads=16,
                                   num_item_tokens_to_predict=32)), cands_c)
        par = topk_parity(cpu_in["got_s"], cpu_in["got_i"], ref_s, ref_i, full, score_tol=SCORE_TOL_BF16,
                          tie_tol=2 * SCORE_TOL_BF16)
        par["rows_of_the_timed_call"] = rows
        cpu_baseline = {"value": Bc / dt, "unit": "users/s", "cores": cores, "kind": kind,
                        "sample": f"{Bc} users of the timed workload (S={Hh * 32} keys, {N} candidates, top-{k}), "
                                  + ("UNMODIFIED reference modules (oracle/_ref)" if kind == "reference" else "oracle port")
                                  + ", fp32 on torch CPU, 1 pass after warm-up",
                        "parity_vs_gpu": par}
        isd, xi = cpu_in["isd"], cpu_in["xi"]
        if ref is not None:
            item_pass = lambda n: RR.item_tokens(r_item, xi[:n], torch.ones(n, 14, dtype=torch.long))
        else:
            item_pass = lambda n: O.item_qformer_forward(isd, xi[:n], None)
        item_pass(2)
        t0 = time.perf_counter()
        item_pass(32)
        items_cpu = {"value": 32 / (time.perf_counter() - t0), "unit": "items/s", "cores": cores, "kind": kind,
                     "sample": "32 items (14 fields x 1024), fp32 on torch CPU"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    def roof(stats, kind, peak, unit_scale):
        if kind not in stats or stats[kind][2] <= 0:
            return None
        n, work, ms = stats[kind]
        ach = work / (ms * 1e-3) / unit_scale
        return {"launches": n, "achieved": ach, "peak": peak, "frac": ach / peak, "avg_launch_ms": ms / n}

    g_user = roof(user_stats, "gemm", pk["bf16_sustained"], 1e12)
    g_item = roof(item_stats, "gemm", pk["bf16_sustained"], 1e12)

    def by_shape(stats, kind, peak, unit_scale):
        rows = {key.split(":", 1)[1]: roof(stats, key, peak, unit_scale) for key in stats if key.startswith(kind + ":")}
        for key, r_ in rows.items():
            r_["share_of_step"] = stats[f"{kind}:{key}"][2] / user_ms
        return dict(sorted(rows.items(), key=lambda kv: -kv[1]["share_of_step"]))

    gemm_shapes = by_shape(user_stats, "gemm", pk["bf16_sustained"], 1e12)
    attn_shapes = by_shape(user_stats, "attention", pk["hbm"], 1e9)
    dom_shape, dom = next(iter(gemm_shapes.items())) if gemm_shapes else (None, None)
    dom_kernel = "gemm_bf16_cg2_kernel<bias>"
    dom_text = (f"gemm_bf16_cg2_kernel<bias>, cross-attention K/V projection, M x N x K = {dom_shape} "
                "(M = users of a K/V chunk x 1600 keys, N = 2 x 1024 per layer x 4 layers [one layer with --layer-major 1]; "
                "tcgen05 cta_group::2, 256 x 256 tiles)")
    dom_flop = dom_alg_bytes = None
    if dom_shape:
        Md, Nd, Kd = (int(x) for x in dom_shape.split("x"))
        dom_flop = 2.0 * Md * Nd * Kd
        dom_alg_bytes = 2 * (Md * Kd + Nd * Kd + Md * Nd)
    kva_shapes = by_shape(user_stats, "kv_attention", pk["bf16_sustained"], 1e12)
    if kva_shapes:
        # --fused-kv 1: the K/V projection lives inside the attention kernel - that launch dominates the step
        kshape, kdom = next(iter(kva_shapes.items()))
        if dom is None or kdom["share_of_step"] > dom["share_of_step"]:
            Mk, Nk, Kk = (int(x) for x in kshape.split("x"))          # rows = users x keys, 2 H, encoder width
            users_l = Mk // (Hh * 32)
            dom_shape, dom = kshape, kdom
            dom_kernel = "kv_attention_umma_kernel" if os.environ.get("UNIREC_KV_ATTENTION_IMPL", "mma_sync") == "umma" \
                else "kv_attention_fused_kernel"
            dom_text = (f"{dom_kernel}: K/V projection of one cross-attention layer (M x N x K = {kshape}, tcgen05 cta_group::2) "
                        "with the 64-query attention over the projected tile fused in (no K/V in HBM); one launch per layer")
            dom_flop = 2.0 * Mk * Nk * Kk + 4.0 * users_l * 16 * 64 * (Hh * 32) * 64
            # sequence read once, weights, queries in, context out, per-split partials written and re-read by the combine
            dom_alg_bytes = 2 * (Mk * Kk + Nk * Kk) + 2 * 2 * users_l * 64 * 1024 + 2 * 4 * users_l * 16 * 4 * (64 * 64 + 128)
    # DRAM traffic of the dominant launch from the committed ncu --set full capture (profiles/, same kernel and shape)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")
    if dom_shape and os.path.exists(tpath):
        tj = json.load(open(tpath))
        for ent in (tj if isinstance(tj, list) else [tj]):
            if ent.get("shape") == dom_shape and ent.get("kernel", "").startswith(dom_kernel.split("<")[0]):
                traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
    out = {
        "metric": METRIC, "value": users_per_sec, "unit": "users/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": user_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_dict(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_users, "unit": "users/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": user_launches,
        "roofline": {
            "kernel": dom_text,
            "bound": "tensor", "achieved": dom["achieved"] if dom else None, "peak": pk["bf16_sustained"],
            "unit": "TFLOP/s", "frac": dom["frac"] if dom else None,
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_flop_per_launch": dom_flop,
            "algorithmic_bytes_per_launch": dom_alg_bytes,
            "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
            "launches": dom["launches"] if dom else 0, "avg_launch_ms": dom["avg_launch_ms"] if dom else None,
            "share_of_step": dom["share_of_step"] if dom else None,
            "end_to_end_frac_of_tensor_peak": users_per_sec / world * FLOPS_PER_USER / (pk["bf16_sustained"] * 1e12),
            "all_gemms": dict(g_user, share_of_step=user_stats["gemm"][2] / user_ms) if g_user else None,
            "gemm_by_shape": gemm_shapes,
            "other_kernels": {
                "attention": roof(user_stats, "attention", pk["hbm"], 1e9),
                "attention_by_shape": attn_shapes,
                "score_topk": roof(user_stats, "score_topk", pk["bf16_sustained"], 1e12),
                "kv_attention_fused": roof(user_stats, "kv_attention", pk["bf16_sustained"], 1e12),
            },
        },
        "cpu_baseline": cpu_baseline,
        "parity_vs_gpu": (cpu_baseline or {}).get("parity_vs_gpu") if world == 1 else multi_parity,
        "torch_eager_same_gpu": eager,
        "kernels_alone": alone,
        "items": {
            "metric": "items/sec (item Q-Former: 14 x 1024 field embeddings -> 32 x 1024 query tokens + pooled row)",
            "value": items_per_sec, "unit": "items/s", "items_timed": items_timed, "ms_total": item_ms,
            "gpu_launches": item_launches,
            "e2e": {"value": Bi * world * n_e2e_items / item_e2e_s, "unit": "items/s",
                    "h2d_bytes_per_step": Bi * 14 * 1024 * 4 + Bi * 14 * 8, "d2h_bytes_per_step": Bi * 32 * 1024 * 2,
                    "how": f"pipeline.generate_item_tokens_streamed over {n_e2e_items} chunks of {Bi} items, pinned host "
                           "fp32 fields in, pinned host bf16 tokens out, copies overlapped with the encode",
                    "serial_copy_encode_copy_items_per_s": Bi * world * 3 / item_e2e_serial_s},
            "roofline": {"bound": "tensor", "achieved": g_item["achieved"] if g_item else None,
                         "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": g_item["frac"] if g_item else None,
                         "share_of_step": (item_stats["gemm"][2] / item_ms) if g_item else None,
                         "end_to_end_frac_of_tensor_peak":
                             items_per_sec / world * FLOPS_PER_ITEM / (pk["bf16_sustained"] * 1e12),
                         "attention": roof(item_stats, "attention", pk["hbm"], 1e9)},
            "cpu_baseline": items_cpu,
        },
        "train": train_block,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
