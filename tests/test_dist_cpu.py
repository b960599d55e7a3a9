"""CPU, world_size 2 over gloo: the host-side multi-GPU logic of config 5 (row-sharded candidate pool,
all-gather of user vectors and of per-rank top-k lists with global indices).  The per-rank top-k here is
computed by the ORACLE (no GPU on this box); what is under test is the sharding + exchange plumbing in
unirec_b200/pipeline.py, which is the same code that runs over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import qformer_oracle as O
        from unirec_b200.pipeline import gather_lists, gather_rows, shard_range
        g = torch.Generator().manual_seed(5)
        B, N, D, k = 6, 1001, 64, 20
        users = torch.randn(B, D, generator=g)
        cands = torch.randn(N, D, generator=g)
        # users are encoded B/G per rank, then all-gathered
        ulo, uhi = shard_range(B, rank, world)
        u_all = gather_rows(users[ulo:uhi].contiguous())
        assert torch.equal(u_all, users)
        # each rank ranks all users against its candidate rows; indices are made global with index_base
        lo, hi = shard_range(N, rank, world)
        s, i = O.cosine_topk(u_all, cands[lo:hi], k)
        s_all, i_all = gather_lists(s, i + lo)
        assert tuple(s_all.shape) == (world, B, k)
        merged_s, pos = torch.topk(s_all.permute(1, 0, 2).reshape(B, world * k), k, dim=-1)
        merged_i = torch.gather(i_all.permute(1, 0, 2).reshape(B, world * k), 1, pos)
        ref_s, ref_i = O.cosine_topk(users, cands, k)
        ok = torch.allclose(merged_s, ref_s, atol=1e-6) and torch.equal(merged_i, ref_i)
        # all-to-all variant: local int32 indices + shard bases on the wire, every rank receives ITS users' lists from
        # every rank and merges only those; one user's list holds -1 entries (a shard with fewer than k candidates)
        from unirec_b200.pipeline import exchange_lists
        bases = gather_rows(torch.tensor([lo], dtype=torch.int64))
        assert bases.tolist() == [shard_range(N, r, world)[0] for r in range(world)]
        i_loc = i.clone()
        s_loc = s.clone()
        i_loc[0, -2:] = -1
        s_loc[0, -2:] = float("-inf")
        s_x, i_x = exchange_lists(s_loc, i_loc, bases)
        b = B // world
        assert tuple(s_x.shape) == (world, b, k) and i_x.dtype == torch.int64
        mine_s, pos = torch.topk(s_x.permute(1, 0, 2).reshape(b, world * k), k, dim=-1)
        mine_i = torch.gather(i_x.permute(1, 0, 2).reshape(b, world * k), 1, pos)
        if rank == 0:
            ok = ok and bool((i_x[:, 0, -2:] == -1).all())           # the marker survives the base shift
            ok = ok and torch.allclose(mine_s[1:], ref_s[ulo + 1:uhi], atol=1e-6) and torch.equal(mine_i[1:], ref_i[ulo + 1:uhi])
        else:
            ok = ok and torch.allclose(mine_s, ref_s[ulo:uhi], atol=1e-6) and torch.equal(mine_i, ref_i[ulo:uhi])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_rank_and_gather_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=5) for _ in range(world))
    assert results == {0: True, 1: True}


def test_gather_is_identity_without_process_group():
    from unirec_b200.pipeline import gather_lists, gather_rows
    s, i = torch.randn(3, 5), torch.arange(15).view(3, 5)
    s_all, i_all = gather_lists(s, i)
    assert tuple(s_all.shape) == (1, 3, 5) and torch.equal(i_all[0], i)
    assert gather_rows(s) is s


def _allreduce_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from unirec_b200.training import GradientAllReducer
        red = GradientAllReducer()
        assert red.world == world
        g = torch.Generator().manual_seed(100 + rank)
        layers = [[torch.randn(7, 5, generator=g), torch.randn(5, generator=g)] for _ in range(3)]
        big = torch.randn(12, 4, generator=g)
        views = [big[:6], big[6:]]                      # gradients handed to autograd as views of a fused buffer
        keep = [[t.clone() for t in ts] for ts in layers] + [[big.clone()]]
        for li in (2, 1, 0):                            # backward order
            red.layer_ready(li, layers[li])
        red.layer_ready(-1, [big, None])
        # in-place bucket: the tensors are views tiling one contiguous buffer (the backbone's flat gradient buffer)
        flatbuf = torch.randn(50, generator=g)
        flat_keep = flatbuf.clone()
        fviews = [flatbuf[:35].view(7, 5), flatbuf[35:40], flatbuf[40:]]
        red.layer_ready(5, fviews, flatbuf)
        red.finish()
        assert not red.pending
        gathered = [torch.empty_like(flat_keep) for _ in range(world)]
        dist.all_gather(gathered, flat_keep)
        assert torch.allclose(flatbuf, sum(gathered) / world, atol=1e-6)
        assert torch.equal(fviews[0].reshape(-1), flatbuf[:35])
        # expected: mean over ranks
        ok = True
        for ts, mine in zip(layers + [[big]], keep):
            for t, m in zip(ts, mine):
                gathered = [torch.empty_like(m) for _ in range(world)]
                dist.all_gather(gathered, m)
                ok = ok and torch.allclose(t, sum(gathered) / world, atol=1e-6)
        ok = ok and torch.equal(torch.cat(views), big)  # views see the in-place result
        # reduce_tensors (behind a CUDA-graph step): views of one buffer are reduced as that buffer, in place; tensors
        # that stand alone go through one packed bucket
        buf = torch.randn(60, generator=g)
        alone = [torch.randn(3, 4, generator=g), torch.randn(5, generator=g)]
        mix = [buf[:24].view(4, 6), buf[24:30], buf[30:].view(3, 10)] + alone
        mix_keep = [t.clone() for t in mix]
        buckets, rest = GradientAllReducer.alias_buckets(mix)
        assert len(buckets) == 1 and buckets[0].numel() == 60 and buckets[0].data_ptr() == buf.data_ptr()
        assert len(rest) == 2
        sparse_cover, lonely = GradientAllReducer.alias_buckets([buf[:5], buf[50:55]])
        assert not sparse_cover and len(lonely) == 2        # two small views far apart: not worth reducing the span
        before = red.bytes_reduced
        red.reduce_tensors(mix + [None])
        assert red.bytes_reduced - before == (60 + 12 + 5) * 4
        for t, m in zip(mix, mix_keep):
            gathered = [torch.empty_like(m) for _ in range(world)]
            dist.all_gather(gathered, m)
            ok = ok and torch.allclose(t, sum(gathered) / world, atol=1e-6)
        # bf16 wire: the fp32 gradient buffer is cast once, reduced as bf16 and written back as fp32 - the mean of the
        # bf16-rounded per-rank values, i.e. within 2^-7 of the mean magnitude (bf16 keeps 8 significant bits: 2^-8 per rounding, inputs and sum)
        red16 = GradientAllReducer(bucket_dtype=torch.bfloat16)
        buf16 = torch.randn(64, generator=g)
        keep16 = buf16.clone()
        lone16 = torch.randn(9, generator=g)
        lone_keep = lone16.clone()
        before = red16.bytes_reduced
        red16.layer_ready(0, [buf16[:40].view(8, 5), buf16[40:]], buf16)
        red16.layer_ready(-3, [lone16])
        red16.finish()
        assert red16.bytes_reduced - before == (64 + 9) * 2 and buf16.dtype == torch.float32
        for t, m in ((buf16, keep16), (lone16, lone_keep)):
            gathered = [torch.empty_like(m) for _ in range(world)]
            dist.all_gather(gathered, m)
            mean = sum(gathered) / world
            bound = 2.0 ** -7 * sum(x.abs() for x in gathered) / world + 1e-6   # one rounding per rank value (2^-8) + one of the sum
            ok = ok and bool(((t - mean).abs() <= bound).all())
        p = torch.nn.Parameter(torch.zeros(3))
        p.grad = torch.full((3,), float(rank + 1))
        red.reduce_params([p, torch.nn.Parameter(torch.zeros(2))])      # second one has no grad: skipped
        ok = ok and torch.allclose(p.grad, torch.full((3,), (1 + world) / 2))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gradient_all_reducer_world2():
    """Bucketed asynchronous gradient averaging of the training step (config 2) over gloo."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_allreduce_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=5) for _ in range(world))
    assert results == {0: True, 1: True}


def test_gradient_all_reducer_single_process_is_noop():
    from unirec_b200.training import GradientAllReducer
    red = GradientAllReducer()
    t = torch.ones(4)
    red.layer_ready(0, [t])
    red.finish()
    assert red.world == 1 and torch.equal(t, torch.ones(4))
