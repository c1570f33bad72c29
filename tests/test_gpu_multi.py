"""GPU, >= 2 devices: row-sharded search with both key-exchange paths (peer memory over NVLink and
NCCL all-gather) must equal the single-GPU result bit for bit.  Skipped on single-GPU boxes."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, exchange, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from rag_arc_b200 import ops, sharded, synth
    n, d, nq, k = 300_001, 128, 200, 50
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev, seed=3)        # same corpus on every rank
    q, _ = synth.dense_queries_cuda(x, nq, seed=4)
    lo, hi = sharded.shard_bounds(n, world, rank)
    idx = sharded.ShardedFlatIndex(x[lo:hi].contiguous(), lo, exchange=exchange)
    ok = True
    for it in range(3):                                                     # exercises both buffer slots
        s, i = idx.search(q, k)
        s_ref, i_ref = ops.dense_topk(x, q, k)
        ok = ok and bool(torch.equal(i, i_ref) and torch.equal(s, s_ref))
    flag = torch.tensor([1 if ok and idx.exchange_used == exchange else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put((int(flag.item()), idx.exchange_used))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_sharded_search_equals_single_gpu(exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, exchange, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    ok, used = out.get()
    assert ok == 1, f"mismatch (exchange used: {used})"


def _owner_worker(rank, world, port, signal, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RAGARC_OWNER_SIGNAL=signal)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from rag_arc_b200 import ops, sharded, synth
    n, d, nq, k = 300_001, 128, 203, 50                                    # nq not divisible by the world size
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev, seed=3)
    q, _ = synth.dense_queries_cuda(x, nq, seed=4)
    lo, hi = sharded.shard_bounds(n, world, rank)
    idx = sharded.ShardedFlatIndex(x[lo:hi].contiguous(), lo)
    s_ref, i_ref = ops.dense_topk(x, q, k)
    qlo, qhi = idx.owned_range(nq)
    ok = True
    for it in range(4):                                                     # both inbox slots, twice
        s, i = idx.search_owned(q, k)
        ok = ok and bool(torch.equal(i, i_ref[qlo:qhi]) and torch.equal(s, s_ref[qlo:qhi]))
    ok = ok and idx.exchange_used.startswith("owner-push") and not idx._owner[(nq, k)].timed_out()
    replay, s, i = idx.capture(q, k, owned=True)
    for it in range(3):
        s, i = replay()
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(i, i_ref[qlo:qhi]) and torch.equal(s, s_ref[qlo:qhi]))
    replay2, finish2, outs2 = idx.capture_owned_overlapped(q, k)               # exchange of call i under scoring of call i+1
    for it in range(5):
        replay2()
    finish2(); torch.cuda.synchronize()
    for s2, i2 in outs2:
        ok = ok and bool(torch.equal(i2, i_ref[qlo:qhi]) and torch.equal(s2, s_ref[qlo:qhi]))
    replay2(); finish2(); torch.cuda.synchronize()                              # leaves both ranks on an even call count
    prepare = lambda q32: ops.normalize_cast(q32, torch.bfloat16, True)
    pipe = sharded.ShardedSearchPipeline(idx, prepare, nq, d, k, owned=True)
    batches = [torch.randn((nq, d), generator=torch.Generator().manual_seed(100 + b)).pin_memory() for b in range(4)]
    prev, results = None, []
    for b in batches:
        t = pipe.submit(b)
        if prev is not None:
            hs, hr = pipe.result(prev); results.append((hs.clone(), hr.clone()))
        prev = t
    hs, hr = pipe.result(prev); results.append((hs.clone(), hr.clone()))
    for b, (hs, hr) in zip(batches, results):
        sr, ir = ops.dense_topk(x, prepare(b.to(dev)), k)
        ok = ok and bool(torch.equal(hr, ir[qlo:qhi].cpu()) and torch.equal(hs, sr[qlo:qhi].cpu()))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("signal", ["0", "1"])
def test_query_owner_exchange_equals_single_gpu(signal):
    """search_owned (merge kernel pushes key rows into the owner's inbox over NVLink, every rank merges
    its own queries): eager, CUDA-graphed and through the host pipeline, against the single-GPU result;
    ordered by a symmetric-memory barrier ("0") or by per-query arrival counters ("1")."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_owner_worker, args=(r, world, port, signal, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get() == 1


def _bm25_worker(rank, world, port, out):
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from rag_arc_b200 import ops, sharded, synth
    from rag_arc_b200.core.retrieval.bm25_index import Bm25Index
    toks, offs = synth.bm25_corpus_tokens(50_001, vocab=8000, seed=13)
    full = Bm25Index.from_token_ids(toks, offs, device=None)
    single = Bm25Index.from_token_ids(toks, offs, device=dev)
    idx = sharded.ShardedBm25Index(full, dev)
    qt, ql = idx.encode_query_ids(synth.bm25_queries_tokens(toks, offs, 64, 8, seed=14))
    s, i = idx.search(qt, ql, 50)
    s_ref, i_ref = ops.bm25_topk(single, qt, ql, 50)
    ok = bool(torch.equal(i, i_ref) and torch.equal(s.view(torch.int64), s_ref.view(torch.int64)))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


def test_sharded_bm25_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_bm25_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get() == 1


def _pipe_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from rag_arc_b200 import ops, sharded, synth
    n, d, nq, k = 200_000, 128, 96, 20
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev, seed=3)
    lo, hi = sharded.shard_bounds(n, world, rank)
    idx = sharded.ShardedFlatIndex(x[lo:hi].contiguous(), lo)
    prepare = lambda q32: ops.normalize_cast(q32, torch.bfloat16, True)
    pipe = sharded.ShardedSearchPipeline(idx, prepare, nq, d, k)
    batches = [torch.randn((nq, d), generator=torch.Generator().manual_seed(100 + i)).pin_memory() for i in range(5)]
    ok, prev = True, None
    results = []
    for b in batches:
        t = pipe.submit(b)
        if prev is not None:
            s, r = pipe.result(prev)
            results.append((s.clone(), r.clone()))
        prev = t
    s, r = pipe.result(prev)
    results.append((s.clone(), r.clone()))
    for b, (s, r) in zip(batches, results):
        s_ref, r_ref = ops.dense_topk(x, prepare(b.to(dev)), k)
        ok = ok and bool(torch.equal(r, r_ref.cpu()) and torch.equal(s, s_ref.cpu()))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


def test_sharded_pipeline_host_in_host_out_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_pipe_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get() == 1


def test_single_process_sharded_index_over_two_gpus():
    """ragarc_sharded_*: one host process, shards on cuda:0 and cuda:1, numpy host buffers."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import numpy as np
    from rag_arc_b200 import synth
    from rag_arc_b200.native_index import NativeFlatIndex, NativeShardedIndex
    X = synth.dense_corpus_np(50_000, 256, seed=8)
    Q, planted = synth.dense_queries_np(X, 64, seed=9)
    flat = NativeFlatIndex(256, "bfloat16", "cosine"); flat.add(X)
    sh = NativeShardedIndex(256, "bfloat16", "cosine", devices=(0, 1, 1)); sh.add(X)
    D, I = flat.search(Q, 20); Ds, Is = sh.search(Q, 20)
    assert np.array_equal(I, Is) and np.array_equal(D, Ds) and (Is[:, 0] == planted).all()
