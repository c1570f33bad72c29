"""GPU: the NCCL form of the multi-GPU search inside the C ABI (``ragarc_comm_*``,
``ragarc_sharded_topk`` = local fused top-k + ncclAllGather + merge), driven through ctypes with raw
device pointers - what a host that is not PyTorch calls.  One rank runs on any box; the two-rank
forms (one process per GPU, and one process with a thread per GPU) need two devices."""
import ctypes
import os
import socket
import threading

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _search(N, comm, shard, lo, q, k, nranks, dev):
    n_local, d = shard.shape
    nq = q.shape[0]
    code = N.BF16
    wsb = int(N.lib.ragarc_sharded_topk_workspace_bytes(n_local, d, code, nq, k, nranks))
    assert wsb > 0
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_sharded_topk(comm, shard.data_ptr(), n_local, d, code, q.data_ptr(), nq, k, lo, s.data_ptr(),
                                          i.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream(dev).cuda_stream),
                "sharded_topk")
        torch.cuda.synchronize(dev)
    return s, i


def test_single_rank_communicator_equals_plain_search(dev):
    from rag_arc_b200 import _native as N
    from rag_arc_b200 import ops, synth
    assert N.lib.ragarc_comm_nccl_version() >= 20000
    x = synth.dense_corpus_cuda(60_000, 128, torch.bfloat16, dev, seed=3)
    q, _ = synth.dense_queries_cuda(x, 70, seed=4)
    uid = ctypes.create_string_buffer(128)
    N.check(N.lib.ragarc_comm_unique_id(uid), "comm_unique_id")
    comm = ctypes.c_void_p()
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_comm_init_rank(uid, 1, 0, ctypes.byref(comm)), "comm_init_rank")
    assert N.lib.ragarc_comm_rank(comm) == 0 and N.lib.ragarc_comm_nranks(comm) == 1
    s, i = _search(N, comm, x, 0, q, 20, 1, dev)
    s_ref, i_ref = ops.dense_topk(x, q, 20)
    assert torch.equal(i, i_ref) and torch.equal(s, s_ref)
    N.lib.ragarc_comm_free(comm)
    with pytest.raises(N.RagArcError):
        N.check(N.lib.ragarc_comm_init_rank(uid, 2, 5, ctypes.byref(comm)), "comm_init_rank")


def _rank_worker(rank, world, uid_bytes, out):
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    from rag_arc_b200 import _native as N
    from rag_arc_b200 import ops, sharded, synth
    comm = ctypes.c_void_p()
    N.check(N.lib.ragarc_comm_init_rank(uid_bytes, world, rank, ctypes.byref(comm)), "comm_init_rank")
    n, d, nq, k = 200_001, 128, 150, 50
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev, seed=3)
    q, _ = synth.dense_queries_cuda(x, nq, seed=4)
    lo, hi = sharded.shard_bounds(n, world, rank)
    ok = True
    for _ in range(2):
        s, i = _search(N, comm, x[lo:hi].contiguous(), lo, q, k, world, dev)
        s_ref, i_ref = ops.dense_topk(x, q, k)
        ok = ok and bool(torch.equal(i, i_ref) and torch.equal(s, s_ref))
    N.lib.ragarc_comm_free(comm)
    out.put((rank, ok))


def test_one_process_per_gpu_unique_id_handshake():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from rag_arc_b200 import _native as N
    uid = ctypes.create_string_buffer(128)
    N.check(N.lib.ragarc_comm_unique_id(uid), "comm_unique_id")
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, uid.raw, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    got = dict(out.get() for _ in range(2))
    assert got == {0: True, 1: True}


def test_one_process_two_gpus_init_all_with_a_thread_per_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from rag_arc_b200 import _native as N
    from rag_arc_b200 import ops, sharded, synth
    comms = (ctypes.c_void_p * 2)()
    N.check(N.lib.ragarc_comm_init_all(2, None, comms), "comm_init_all")
    n, d, nq, k = 100_000, 64, 40, 10
    devs = [torch.device("cuda", g) for g in range(2)]
    xs = [synth.dense_corpus_cuda(n, d, torch.bfloat16, dv, seed=3) for dv in devs]
    qs = [synth.dense_queries_cuda(xs[g], nq, seed=4)[0] for g in range(2)]
    res = [None, None]

    def work(g):
        lo, hi = sharded.shard_bounds(n, 2, g)
        res[g] = _search(N, ctypes.c_void_p(comms[g]), xs[g][lo:hi].contiguous(), lo, qs[g], k, 2, devs[g])

    ts = [threading.Thread(target=work, args=(g,)) for g in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    s_ref, i_ref = ops.dense_topk(xs[0], qs[0], k)
    for g in range(2):
        assert torch.equal(res[g][1].cpu(), i_ref.cpu()) and torch.equal(res[g][0].cpu(), s_ref.cpu())
        N.lib.ragarc_comm_free(ctypes.c_void_p(comms[g]))
