"""Two ranks over NCCL (needs 2 GPUs; skipped otherwise): the data-parallel training step of BASELINE config 2
(training/item_qformer_training.py:117-131 split over ranks).  Gradients of the per-rank half batches, averaged by
`GradientAllReducer` - eagerly with the per-layer buckets, and INSIDE the captured CUDA graph
(`TrainStepGraph(reducer=...)`) - equal the single-GPU gradients of the concatenated batch; and the merged top-k of the
row-sharded candidate pool equals the single-GPU list (config 5)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _leave(q):
    """End a worker without tearing NCCL down: destroy_process_group() blocks for minutes while a captured CUDA graph still
    references the communicator, and a test process has nothing to release gracefully."""
    import sys
    sys.stdout.flush()
    sys.stderr.flush()
    q.close()
    q.join_thread()
    os._exit(0)


def _grad_worker(rank, world, port, wire, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from tests.golden_cases import ITEM_CASES
        from unirec_b200 import synth
        from unirec_b200.modules import QFormerForItemRepresentation
        from unirec_b200.training import GradientAllReducer, TrainStepGraph, qformer_loss
        c = ITEM_CASES["small"]
        mk = c["model"]

        def make():
            m = QFormerForItemRepresentation(hidden_size=mk["hidden"], num_hidden_layers=mk["layers"],
                                             num_attention_heads=c["heads"], intermediate_size=mk["inter"],
                                             num_query_tokens=mk["num_query"], field_embedding_dim=mk["field_dim"],
                                             num_fields=mk["num_fields"], dropout=0.0)
            m.load_state_dict(synth.item_qformer_state_dict(**mk, seed=61, attn_std=0.1), strict=True)
            return m.to(dev).train()

        B = 64
        x, m = synth.item_fields(batch=B, num_fields=mk["num_fields"], dim=mk["field_dim"], seed=90, clip_field=2, presence=1.0)
        g = torch.Generator().manual_seed(7)
        pos, neg = torch.randn(B, mk["hidden"], generator=g), torch.randn(B, mk["hidden"], generator=g)
        x, m, pos, neg = x.to(dev), m.to(dev), pos.to(dev), neg.to(dev)
        lo, hi = rank * B // world, (rank + 1) * B // world

        # single-GPU gradients of the whole batch (every rank computes them; equal halves + all-ones masks make the
        # mean of the per-rank losses the loss of the concatenated batch)
        full = make()
        qformer_loss(full(x, m), x, m, pos, neg).backward()
        ref = {n: p.grad.clone() for n, p in full.named_parameters() if p.grad is not None}
        del full

        def compare(model, what):
            worst_cos, worst_rel = 1.0, 0.0
            for n, p in model.named_parameters():
                if p.grad is None:
                    assert n not in ref, n
                    continue
                a, b = p.grad.float().flatten(), ref[n].float().flatten()
                if float(b.norm()) < 1e-6:
                    continue
                cos = float(torch.nn.functional.cosine_similarity(a, b, dim=0))
                rel = float((a - b).norm() / b.norm())
                worst_cos, worst_rel = min(worst_cos, cos), max(worst_rel, rel)
            print(f"rank {rank} {what} wire={wire}: worst cos {worst_cos:.6f} worst rel {worst_rel:.4f}", flush=True)
            # bf16 kernels on half batches vs the full batch: different tile / split-K order; bf16 wire adds 2^-8 per value
            return worst_cos >= 0.9995 and worst_rel <= (0.03 if wire == torch.bfloat16 else 0.025)

        ok = True
        # (1) eager step: per-layer buckets start their all-reduce from inside the backbone's backward
        model = make()
        red = GradientAllReducer(bucket_dtype=wire).attach(model.qformer)
        heads = [p for n, p in model.named_parameters() if not n.startswith("qformer.") and n != "query_embeddings"]
        qformer_loss(model(x[lo:hi], m[lo:hi]), x[lo:hi], m[lo:hi], pos[lo:hi], neg[lo:hi]).backward()
        red.reduce_params(heads)
        ok = ok and compare(model, "eager")
        model.zero_grad(set_to_none=True)
        import gc
        gc.collect()
        # (2) the same step with the collectives captured inside the CUDA graph; two replays give the same gradients
        tg = TrainStepGraph(model, x[lo:hi], m[lo:hi], reducer=red)
        assert tg.allreduce_bytes_per_step > 0
        tg.step(x[lo:hi], m[lo:hi], pos[lo:hi], neg[lo:hi])
        ok = ok and compare(model, "graph replay 1")
        first = [g_.clone() for g_ in tg.grad_tensors()]
        tg.step(x[lo:hi], m[lo:hi], pos[lo:hi], neg[lo:hi])
        ok = ok and compare(model, "graph replay 2")
        same = all(float((a - b).abs().max()) <= 1e-3 * float(a.abs().max()) + 1e-7 for a, b in zip(first, tg.grad_tensors()))
        torch.cuda.synchronize()
        q.put((rank, bool(ok and same)))
    except Exception:
        import traceback
        traceback.print_exc()
        q.put((rank, False))
    _leave(q)


def _rank_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from unirec_b200 import ops, synth
        from unirec_b200.modules import UserQFormer
        from unirec_b200.pipeline import NestedRanker, shard_range
        N, Q, D, B, Hmax, k = 3000, 32, 256, 12, 5, 20
        table = synth.normal("tok_table_dp", (N, Q, D), 35).to(torch.bfloat16).to(dev)
        cands = table.float().mean(dim=1).to(torch.bfloat16)
        mkw = dict(hidden=256, layers=2, inter=512, num_query=64, input_dim=256, num_predict=8)
        um = UserQFormer(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                         num_query_tokens=64, input_embedding_dim=256, num_item_tokens_to_predict=8)
        um.load_state_dict(synth.user_qformer_state_dict(**mkw, seed=31, attn_std=0.1), strict=True)
        um = um.to(dev).eval()
        gen = torch.Generator().manual_seed(36)
        history = torch.randint(0, N, (B, Hmax), generator=gen).to(dev)
        lengths = torch.randint(1, Hmax + 1, (B,), generator=gen, dtype=torch.int32).to(dev)
        # single GPU: all users, whole pool - on a one-rank group (group=None would be the 2-rank default group)
        solo = [dist.new_group([r_]) for r_ in range(world)][rank]
        one = NestedRanker(um, table, cands, k=k, group=solo)
        u_all = one.encode_users(history, lengths)
        s1, i1 = one.rank(u_all)
        # two ranks: users split, candidate rows split, all-gather of the user vectors, per-rank top-k with global indices,
        # all-gather of the lists + merge.  Same user vectors in -> the SAME list out (scores bit-equal: a dot product does
        # not depend on the tile it is computed in; ties broken by the smaller global index on both paths)
        lo, hi = shard_range(N, rank, world)
        ulo, uhi = shard_range(B, rank, world)
        two = NestedRanker(um, table, cands[lo:hi].contiguous(), k=k, index_base=lo, group=dist.group.WORLD)
        s2, i2 = two.rank(u_all[ulo:uhi].contiguous())
        ok = tuple(s2.shape) == (B, k) and bool(torch.equal(s1, s2)) and bool(torch.equal(i1, i2))
        # and end to end (this rank encodes only its users): the same vectors up to the batch composition of the kernels
        u_mine = two.encode_users(history[ulo:uhi].contiguous(), lengths[ulo:uhi].contiguous())
        cos = torch.nn.functional.cosine_similarity(u_mine.float(), u_all[ulo:uhi].float(), dim=-1)
        ok = ok and float(cos.min()) > 0.9999
        s3, i3 = two(history[ulo:uhi].contiguous(), lengths[ulo:uhi].contiguous())
        ok = ok and bool(torch.allclose(s1, s3, atol=2e-3))
        # all-to-all exchange: this rank's users only, the same lists
        s4, i4 = two.rank(u_all[ulo:uhi].contiguous(), local_result=True)
        ok = ok and bool(torch.equal(s4, s1[ulo:uhi])) and bool(torch.equal(i4, i1[ulo:uhi]))
        print(f"rank {rank}: merged == single-GPU list: {bool(torch.equal(i1, i2))}, min cos {float(cos.min()):.6f}", flush=True)
        if not ok:
            print(f"rank {rank}: scores equal {bool(torch.equal(s1, s2))} (max diff {float((s1 - s2).abs().max()):.3g}), "
                  f"{int((i1 != i2).sum())} of {i1.numel()} indices differ; first rows:\n{i1[0, :8].tolist()}\n{i2[0, :8].tolist()}\n"
                  f"{s1[0, :8].tolist()}\n{s2[0, :8].tolist()}", flush=True)
        torch.cuda.synchronize()
        q.put((rank, bool(ok)))
    except Exception:
        import traceback
        traceback.print_exc()
        q.put((rank, False))
    _leave(q)


def _run(target, *args):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        results = dict(q.get(timeout=240) for _ in range(world))
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()                     # our own child, by handle
    assert results == {0: True, 1: True}


needs2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")


@needs2
@pytest.mark.parametrize("wire", [torch.float32, torch.bfloat16])
def test_data_parallel_gradients_equal_single_gpu_gradients(wire):
    _run(_grad_worker, wire)


@needs2
def test_sharded_ranking_equals_single_gpu_ranking():
    _run(_rank_worker)
