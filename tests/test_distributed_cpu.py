"""CPU, world_size 2, gloo: the multi-rank plumbing of ``ShardedFlatIndex`` (shard bounds, global
ids packed into the keys, one all-gather, merge on every rank).  The two CUDA entry points are
replaced by oracle-backed stand-ins - the kernels themselves are covered by the GPU tests
(tests/test_gpu_dense.py::test_keys_and_merge_equal_single_shot)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _ord(f32):
    u = f32.astype(np.float32).view(np.uint32)
    return np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint64)


def _fake_topk_keys(corpus, queries, k, id_base, n_rows=None, **kw):
    from oracle import dense as odense
    X = corpus[:n_rows].float().numpy(); Q = queries.float().numpy()
    D, I = odense.flat_ip_search(X, Q, k)
    keys = (_ord(D) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - (I.astype(np.uint64) + np.uint64(id_base)))
    keys[I < 0] = 0
    return torch.from_numpy(keys.view(np.int64))


def _fake_merge(keys, k_out):
    G, nq, k = keys.shape
    flat = keys.numpy().view(np.uint64).transpose(1, 0, 2).reshape(nq, G * k)
    srt = np.sort(flat, axis=1)[:, ::-1][:, :k_out]            # larger key = better (score, then lower id)
    ids = (np.uint64(0xFFFFFFFF) - (srt & np.uint64(0xFFFFFFFF))).astype(np.int64)
    o = (srt >> np.uint64(32)).astype(np.uint32)
    u = np.where(o & 0x80000000, o & 0x7FFFFFFF, ~o).astype(np.uint32)
    scores = u.view(np.float32).copy()
    ids[srt == 0] = -1
    return torch.from_numpy(scores), torch.from_numpy(ids)


def _worker(rank, world, port, n, d, nq, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rag_arc_b200 import ops, sharded
    ops.dense_topk_keys = _fake_topk_keys
    ops.merge_topk_keys = _fake_merge
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    lo, hi = sharded.shard_bounds(n, world, rank)
    idx = sharded.ShardedFlatIndex(torch.from_numpy(X[lo:hi]), lo)
    scores, ids = idx.search(torch.from_numpy(Q), k)
    from oracle import dense as odense
    D, I = odense.flat_ip_search(X, Q, k)
    ok = bool((ids.numpy() == I).all() and np.array_equal(scores.numpy(), D))
    # query-owner form: without peer memory (CPU, gloo) it falls back to the all-gather + slice; the owned
    # ranges of the ranks partition the batch and every rank holds exactly its slice of the global result
    qlo, qhi = idx.owned_range(nq)
    s_own, i_own = idx.search_owned(torch.from_numpy(Q), k)
    ok = ok and bool((i_own.numpy() == I[qlo:qhi]).all() and np.array_equal(s_own.numpy(), D[qlo:qhi]))
    bounds = [None] * world
    dist.all_gather_object(bounds, (qlo, qhi))
    ok = ok and bounds[0][0] == 0 and bounds[-1][1] == nq and all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n,k", [(1001, 10), (7, 5)])
def test_sharded_search_two_ranks_gloo_equals_single_shot(n, k):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 16, 5, k, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() == 1


# ---- doc-range sharded BM25 ----------------------------------------------------------------------
def _fake_bm25_topk(index, q_terms, q_len, k, use_post_val=True):
    """numpy stand-in for ragarc_bm25_topk over the shard's own arrays (reference operation order)."""
    from oracle import bm25 as obm25
    qt = q_terms.numpy(); ql = q_len.numpy()
    nq = qt.shape[0]
    S = np.full((nq, k), -np.inf); I = np.full((nq, k), -1, np.int64)
    for q in range(nq):
        sc = np.zeros(index.n_docs)
        for t in qt[q, :ql[q]]:
            if t < 0:
                continue
            a, b = index.indptr_np[t], index.indptr_np[t + 1]
            sc[index.post_doc_np[a:b]] += index.idf_np[t] * index.post_val_np[a:b]
        top = obm25.stable_topk(sc, min(k, index.n_docs))
        S[q, :len(top)] = sc[top]; I[q, :len(top)] = top + index.id_base
    return torch.from_numpy(S), torch.from_numpy(I)


def _fake_bm25_merge(scores, ids, k_out):
    G, nq, k = scores.shape
    s = scores.numpy().transpose(1, 0, 2).reshape(nq, G * k)
    i = ids.numpy().transpose(1, 0, 2).reshape(nq, G * k)
    out_s = np.full((nq, k_out), -np.inf); out_i = np.full((nq, k_out), -1, np.int64)
    for q in range(nq):
        valid = np.flatnonzero(i[q] >= 0)
        order = valid[np.lexsort((i[q][valid], -s[q][valid]))][:k_out]
        out_s[q, :len(order)] = s[q][order]; out_i[q, :len(order)] = i[q][order]
    return torch.from_numpy(out_s), torch.from_numpy(out_i)


def _bm25_worker(rank, world, port, n_docs, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bm25 as obm25
    from rag_arc_b200 import ops, sharded
    from rag_arc_b200.core.retrieval.bm25_index import Bm25Index
    ops.bm25_topk = _fake_bm25_topk
    ops.bm25_merge_topk = _fake_bm25_merge
    rng = np.random.default_rng(3)
    words = [f"w{i}" for i in range(40)]
    corpus = [[words[j] for j in rng.integers(0, 40, rng.integers(1, 12))] for _ in range(n_docs)]
    queries = [[words[j] for j in rng.integers(0, 40, 4)] + ["unseen"] for _ in range(6)]
    full = Bm25Index.from_token_lists(corpus, device=None)
    idx = sharded.ShardedBm25Index(full, "cpu")
    qt, ql = idx.encode_queries(queries)
    scores, ids = idx.search(qt, ql, k)
    ref = obm25.BM25Okapi(corpus)
    ok = True
    for qi, q in enumerate(queries):
        want = ref.get_scores(q)
        top = obm25.stable_topk(want, min(k, n_docs))
        ok &= ids[qi, :len(top)].tolist() == top.tolist()
        ok &= np.array_equal(scores[qi, :len(top)].numpy().view(np.uint64), want[top].view(np.uint64))
        ok &= bool((ids[qi, len(top):] == -1).all())
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_docs,k", [(101, 10), (3, 5)])
def test_sharded_bm25_two_ranks_gloo_equals_single_index(n_docs, k):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_bm25_worker, args=(r, 2, port, n_docs, k, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() == 1
