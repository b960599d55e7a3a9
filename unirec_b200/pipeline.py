"""Host-side drivers of the nested encode-and-rank path.

  * generate_item_tokens  - batched item-query-token generation, item-range sharded (config 3;
                            intent of data_processing/generate_all_item_embeddings.py:238-267, working
                            equivalent data_processing/qformer_inference.py:143-168): no collective.
  * generate_item_tokens_streamed - the same loop from pinned host memory to pinned host memory with the two
                            transfers on their own streams (the reference's copy -> model -> .cpu() pattern).
  * NestedRanker          - user-sequence build (models/user_sequence_encoder.py:128-140) -> UserQFormer
                            -> pooled scoring vector -> cosine top-k over a row-sharded candidate pool with
                            an NCCL all-gather + merge of the per-GPU top-k lists (config 5).
One process per GPU; torch.distributed is used for the user-vector all-gather and the exchange of the per-rank top-k
lists only (all-gather: every rank ends with every user's list; all-to-all: every rank merges just its own users).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple, Union

import torch

from . import ops
from .modules import QFormerForItemRepresentation, UserQFormer


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank` when n rows are split over world_size ranks
    (ranges differ by at most one row)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


FieldSource = Union[torch.Tensor, Callable[[int, int], Tuple[torch.Tensor, Optional[torch.Tensor]]]]


@torch.no_grad()
def generate_item_tokens(model: QFormerForItemRepresentation, fields: FieldSource, num_items: Optional[int] = None,
                         mask: Optional[torch.Tensor] = None, *, batch_size: int = 4096, rank: int = 0,
                         world_size: int = 1, tokens_out: Optional[torch.Tensor] = None,
                         pooled_out: Optional[torch.Tensor] = None, keep_tokens: bool = True):
    """Run the item Q-Former over this rank's item range in chunks of `batch_size`.

    fields: a [N, F, E] tensor (CUDA) or a callable (lo, hi) -> (fields[hi-lo, F, E], mask or None)
    producing the chunk on the device.  Returns (tokens bf16 [n_local, Q, H] or None, pooled bf16
    [n_local, H], (lo, hi)); `pooled` = tokens.mean(dim=1), the candidate scoring vector.
    """
    if isinstance(fields, torch.Tensor):
        n_total = fields.shape[0]
    else:
        if num_items is None:
            raise ValueError("num_items is required when fields is a callable")
        n_total = num_items
    lo, hi = shard_range(n_total, rank, world_size)
    n_local = hi - lo
    Q, H = model.num_query_tokens, model.config.hidden_size
    dev = next(model.parameters()).device
    if keep_tokens and tokens_out is None:
        tokens_out = torch.empty(n_local, Q, H, device=dev, dtype=torch.bfloat16)
    if pooled_out is None:
        pooled_out = torch.empty(n_local, H, device=dev, dtype=torch.bfloat16)
    for c0 in range(lo, hi, batch_size):
        c1 = min(c0 + batch_size, hi)
        if isinstance(fields, torch.Tensor):
            x, m = fields[c0:c1], (None if mask is None else mask[c0:c1])
        else:
            x, m = fields(c0, c1)
        tok = model.encode_query_tokens(x, m, out_dtype=torch.bfloat16)
        if keep_tokens:
            tokens_out[c0 - lo:c1 - lo].copy_(tok)
        pooled_out[c0 - lo:c1 - lo].copy_(ops.mean_tokens(tok))
    return (tokens_out if keep_tokens else None), pooled_out, (lo, hi)


@torch.no_grad()
def generate_item_tokens_streamed(model: QFormerForItemRepresentation, fields_host: torch.Tensor,
                                  mask_host: Optional[torch.Tensor], tokens_host_out: torch.Tensor, *,
                                  batch_size: int = 4096, pooled_out: Optional[torch.Tensor] = None, depth: int = 2,
                                  device: Optional[torch.device] = None) -> torch.Tensor:
    """Host-to-host item-token generation - the I/O pattern of the reference loop (data_processing/
    qformer_inference.py:143-168: a host batch goes to the device, `query_outputs` comes back with `.cpu()`), with the
    transfers taken off the critical path: the host-to-device copy of chunk i+1 and the device-to-host copy of chunk i-1
    run on their own streams while chunk i is encoded (the reference serialises copy -> model -> copy and synchronises
    per batch).  fields_host [N, F, E] fp32 and tokens_host_out [N, Q, H] bf16 should be PINNED for the copies to be
    asynchronous; at most `depth` chunks are in flight.  Returns pooled candidate vectors bf16 [N, H] on the device
    (`pooled_out` if given).  On return `tokens_host_out` is COMPLETE: the last device-to-host copy has been waited for on
    the host (the reference's `.cpu()` is host-synchronous too); `pooled_out` is ordered on the caller's stream."""
    dev = device or next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("generate_item_tokens_streamed: the model must live on a CUDA device (no CPU path)")
    n = fields_host.shape[0]
    H = model.config.hidden_size
    if pooled_out is None:
        pooled_out = torch.empty(n, H, device=dev, dtype=torch.bfloat16)
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    done = []                                            # per chunk: event "encoded and copied out"
    for i, c0 in enumerate(range(0, n, batch_size)):
        c1 = min(c0 + batch_size, n)
        if i >= depth:
            done[i - depth].synchronize()                # bounds the device memory held by chunks in flight
        with torch.cuda.stream(s_in):
            x = fields_host[c0:c1].to(dev, non_blocking=True)
            m = None if mask_host is None else mask_host[c0:c1].to(dev, non_blocking=True)
            ev_in = s_in.record_event()
        cur.wait_event(ev_in)
        x.record_stream(cur)
        if m is not None:
            m.record_stream(cur)
        tok = model.encode_query_tokens(x, m, out_dtype=torch.bfloat16)
        pooled_out[c0:c1].copy_(ops.mean_tokens(tok))
        ev_c = cur.record_event()
        s_out.wait_event(ev_c)
        with torch.cuda.stream(s_out):
            tokens_host_out[c0:c1].copy_(tok, non_blocking=True)
            done.append(s_out.record_event())
        tok.record_stream(s_out)
    cur.wait_stream(s_out)
    if done:
        done[-1].synchronize()          # copies on s_out complete in order: the host buffer is fully written
    return pooled_out


def gather_lists(scores: torch.Tensor, idx: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """all-gather per-rank top-k lists [B, k] -> [G, B, k] (works on NCCL and gloo)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return scores.unsqueeze(0), idx.unsqueeze(0)
    g = dist.get_world_size(group)
    b = scores.shape[0]
    # concatenated-along-dim-0 output layout is the one every backend (NCCL, gloo) accepts
    s_all = torch.empty((g * b,) + tuple(scores.shape[1:]), dtype=scores.dtype, device=scores.device)
    i_all = torch.empty((g * b,) + tuple(idx.shape[1:]), dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(s_all, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(i_all, idx.contiguous(), group=group)
    return s_all.view((g,) + tuple(scores.shape)), i_all.view((g,) + tuple(idx.shape))


def exchange_lists(scores: torch.Tensor, idx_local: torch.Tensor, bases: torch.Tensor, group=None
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """all-to-all of per-rank top-k lists: this rank holds the lists of ALL G * b users against ITS candidate rows
    (scores fp32 [G * b, k], idx_local = row numbers inside this rank's shard, [G * b, k]); it sends block j (the users
    rank j encoded) to rank j and receives its own b users' lists from every rank -> (scores fp32 [G, b, k], global idx
    int64 [G, b, k]), list g computed against rank g's rows.  `bases` int64 [G]: global row number of every rank's first
    candidate.  One collective: scores and indices travel together as int32 pairs (8 bytes per entry instead of the 12
    of an fp32 + int64 all-gather), and each rank merges b users instead of G * b."""
    import torch.distributed as dist
    g = dist.get_world_size(group)
    n, k = scores.shape
    if n % g != 0:
        raise ValueError(f"exchange_lists: {n} users do not split over {g} ranks")
    b = n // g
    send = torch.stack([scores.contiguous().view(torch.int32), idx_local.to(torch.int32)], dim=1).contiguous()   # [G*b, 2, k]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    recv = recv.view(g, b, 2, k)
    s = recv[:, :, 0].contiguous().view(torch.float32)
    i = recv[:, :, 1].to(torch.int64)
    i = torch.where(i < 0, i, i + bases.to(recv.device).view(g, 1, 1))       # -1 marks "fewer than k candidates"
    return s, i


def gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather equal-sized row blocks [b, D] -> [G*b, D]."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    g = dist.get_world_size(group)
    out = torch.empty((g * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


class NestedRanker:
    """item-token table -> user sequence -> UserQFormer -> pooled vector -> top-k over the candidate pool.

    item_tokens : bf16 [N_items, Q, D]  item query-token table used for history gathering (replicated)
    candidates  : bf16 [N_local, D]     this rank's rows of the pooled candidate table
    index_base  : global row index of candidates[0]
    """

    def __init__(self, user_model: UserQFormer, item_tokens: torch.Tensor, candidates: torch.Tensor, k: int = 100,
                 index_base: int = 0, group=None, fused_gather: bool = False):
        self.user_model = user_model
        self.item_tokens = item_tokens
        self.candidates = candidates
        self.cand_inv = ops.inv_l2_norm(candidates)   # candidate norms are computed once, not per query batch
        self.k = k
        self.index_base = index_base
        self.group = group
        # True: the K/V projection gathers the history tokens itself (UserQFormer.encode_queries_from_history) instead of
        # reading a materialised user sequence; used when no per-user context vector is given
        self.fused_gather = fused_gather
        self.last_user_vectors: Optional[torch.Tensor] = None
        self._bases: Optional[torch.Tensor] = None

    @torch.no_grad()
    def encode_users(self, history: torch.Tensor, lengths: torch.Tensor,
                     context: Optional[torch.Tensor] = None) -> torch.Tensor:
        """history int64 [B, Hmax], lengths int32 [B] -> pooled predicted-item vector bf16 [B, D]."""
        um = self.user_model
        B = history.shape[0]
        S = history.shape[1] * self.item_tokens.shape[1]
        if self.fused_gather and context is None:
            hidden = um.encode_queries_from_history(self.item_tokens, history, lengths, torch.bfloat16)
            return ops.mean_tokens(um.predict_from_queries(hidden, out_dtype=torch.bfloat16))
        step = um._chunk_users(S)
        outs = []
        for lo in range(0, B, step):
            hi = min(lo + step, B)
            seq, mask = ops.build_user_sequence(self.item_tokens, history[lo:hi].contiguous(),
                                                lengths[lo:hi].contiguous(),
                                                None if context is None else context[lo:hi])
            hidden = um.encode_queries(seq, mask, torch.bfloat16)
            pred = um.predict_from_queries(hidden, out_dtype=torch.bfloat16)     # [b, T, D]
            outs.append(ops.mean_tokens(pred))                                   # scoring vector
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def _shard_bases(self) -> torch.Tensor:
        """Global row number of every rank's first candidate (int64 [G], gathered once)."""
        if self._bases is None:
            import torch.distributed as dist
            mine = torch.tensor([self.index_base], dtype=torch.int64, device=self.candidates.device)
            self._bases = gather_rows(mine, self.group)
        return self._bases

    @torch.no_grad()
    def rank(self, user_vectors: torch.Tensor, local_result: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """user_vectors bf16 [B_local, D] (this rank's users).  Returns (scores fp32 [B, k], idx int64 [B, k]):
        local_result = False - for ALL users of the group, identical on every rank (all-gather of the per-rank lists,
        every rank merges every user); local_result = True - for THIS rank's B_local users only (all-to-all: a rank
        receives its users' lists from every rank and merges just those - 1/G of the merge work, of the bytes it receives
        and of the result it hands back to the host)."""
        u_all = gather_rows(user_vectors, self.group)
        self.last_user_vectors = u_all          # kept for parity checks on rows of a timed call (bench.py, tests)
        world = 1 if u_all is user_vectors else u_all.shape[0] // user_vectors.shape[0]
        if world == 1:
            return ops.score_topk(u_all, self.candidates, self.k, cand_inv=self.cand_inv, index_base=self.index_base)
        if local_result:
            s, i = ops.score_topk(u_all, self.candidates, self.k, cand_inv=self.cand_inv, index_base=0)
            s_all, i_all = exchange_lists(s, i, self._shard_bases(), self.group)
        else:
            s, i = ops.score_topk(u_all, self.candidates, self.k, cand_inv=self.cand_inv, index_base=self.index_base)
            s_all, i_all = gather_lists(s, i, self.group)
        return ops.topk_merge(s_all, i_all)

    def __call__(self, history, lengths, context=None, local_result: bool = False):
        return self.rank(self.encode_users(history, lengths, context), local_result)
