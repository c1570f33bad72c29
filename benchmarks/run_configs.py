#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configurations other than the bench.py headline
(C3).  One JSON line per measurement on stdout.  Single GPU; the multi-GPU configurations are
measured as the per-GPU shard they put on one device (flat search is exactly linear in rows and the
merge traffic is < 1 MB per rank).

  C1  dense fp32 10k x 384, 100 queries, top-10           (SIMT fp32 path, the reference's precision)
  C2  hybrid: dense 100k x 768 bf16 + BM25 100k docs, top-50 each, RRF, batch 256
  C4s per-GPU shard of C4 at G=8: 1.25M x 1024 fp16, batch 1024, top-100
  C5s per-GPU shard of C5 at G=8: 6.25M x 768 bf16, batch 1..64, top-100  (HBM-bound regime)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from rag_arc_b200 import _native as N  # noqa: E402
from rag_arc_b200 import ops, synth  # noqa: E402
from rag_arc_b200.core.retrieval.bm25_index import Bm25Index  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(iters):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def graph_time(fn, iters=50):
    """Device time per call with the launch overhead taken out: the call is captured in a CUDA graph
    (after a warm-up on the capture stream, so that workspaces exist) and replayed back to back."""
    dev = torch.device("cuda:0")
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn(); fn()
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(**kw):
    print(json.dumps(kw), flush=True)


def dense_case(name, n, d, dtype, nq, k, dev, iters=20):
    if dtype == torch.float32:
        x = torch.from_numpy(synth.dense_corpus_np(n, d)).to(dev)
    else:
        x = synth.dense_corpus_cuda(n, d, dtype, dev)
    q, planted = synth.dense_queries_cuda(x, nq)
    N.profile_enable(True); N.profile_read()
    med, best = timeit(lambda: ops.dense_topk(x, q, k), iters)
    seed_ms, score_ms, merge_ms, nrec = N.profile_read()
    N.profile_enable(False)
    s, i = ops.dense_topk(x, q, k)
    assert bool((i[:, 0] == planted).all())
    flops = 2.0 * nq * n * d
    bytes_ = n * d * x.element_size()
    kern = score_ms / nrec
    emit(config=name, rows=n, dim=d, dtype=str(dtype).replace("torch.", ""), batch=nq, k=k,
         ms_median=med, ms_best=best, qps=nq / (med * 1e-3),
         kernel_ms=kern, seed_ms=seed_ms / nrec, merge_ms=merge_ms / nrec,
         tflops=flops / (kern * 1e-3) / 1e12, frac_tensor_peak=flops / (kern * 1e-3) / 1e12 / PEAKS["bf16_tflops"],
         hbm_gbs=bytes_ / (kern * 1e-3) / 1e9, frac_hbm_peak=bytes_ / (kern * 1e-3) / 1e9 / PEAKS["hbm_gbs"],
         bound="hbm" if bytes_ / (PEAKS["hbm_gbs"] * 1e9) > flops / (PEAKS["bf16_tflops"] * 1e12) else "tensor")
    del x, q
    torch.cuda.empty_cache()


def x3_case(dev):
    """fp32 storage, 100k x 768, batch 256: CUDA-core fp32 path vs bf16x3 tensor-core path."""
    n, d, nq, k = 100_000, 768, 256, 50
    X = torch.from_numpy(synth.dense_corpus_np(n, d)).to(dev)
    q, _ = synth.dense_queries_cuda(X, nq)
    xp = ops.normalize_split3(X, False); qp = ops.normalize_split3(q, False)
    m_simt, _ = timeit(lambda: ops.dense_topk(X, q, k), 5)
    m_x3, _ = timeit(lambda: ops.dense_topk_x3(xp, qp, k), 20)
    s1, i1 = ops.dense_topk(X, q, k); s2, i2 = ops.dense_topk_x3(xp, qp, k)
    emit(config="fp32-100kx768-Q256", ms_simt=m_simt, ms_bf16x3=m_x3, speedup=m_simt / m_x3,
         ids_equal=float((i1 == i2).float().mean()), max_abs_score_diff=float((s1 - s2).abs().max()),
         eff_tflops_x3=2.0 * nq * n * d / (m_x3 * 1e-3) / 1e12)


def c2_hybrid(dev):
    n, d, nq, k = 100_000, 768, 256, 50
    toks, offs = synth.bm25_corpus_tokens(n)
    qtok = synth.bm25_queries_tokens(toks, offs, nq)
    t0 = time.perf_counter()
    idx = Bm25Index.from_token_ids(toks, offs, device=dev)
    build_s = time.perf_counter() - t0
    qt, ql = idx.encode_query_ids(qtok)
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev)
    q, _ = synth.dense_queries_cuda(x, nq)
    # algorithmic bytes of the BM25 scoring: postings 8 B each + touched 16 B + N*8 zero/read
    df = np.diff(idx.indptr_np)
    qterms = qt.cpu().numpy()
    sum_df = int(sum(df[t] for row in qterms for t in row if t >= 0))
    bm_med, bm_best = timeit(lambda: ops.bm25_topk(idx, qt, ql, k), 10)
    bytes_bm = sum_df * 8 + sum_df * 16 + nq * n * 8 * 2
    bm_graph = graph_time(lambda: ops.bm25_topk(idx, qt, ql, k))
    emit(config="C2-bm25", docs=n, batch=nq, k=k, ms_median=bm_med, ms_best=bm_best, ms_graph=bm_graph,
         qps=nq / (bm_med * 1e-3), qps_graph=nq / (bm_graph * 1e-3),
         postings_per_query=sum_df / nq, algorithmic_bytes=bytes_bm, hbm_gbs=bytes_bm / (bm_med * 1e-3) / 1e9,
         frac_hbm_peak=bytes_bm / (bm_med * 1e-3) / 1e9 / PEAKS["hbm_gbs"], index_build_s=build_s, nnz=int(df.sum()))
    de_med, de_best = timeit(lambda: ops.dense_topk(x, q, k), 20)
    de_graph = graph_time(lambda: ops.dense_topk(x, q, k))
    emit(config="C2-dense", rows=n, dim=d, batch=nq, k=k, ms_median=de_med, ms_best=de_best, ms_graph=de_graph,
         qps=nq / (de_med * 1e-3))

    def hybrid():
        _, bid = ops.bm25_topk(idx, qt, ql, k)
        _, did = ops.dense_topk(x, q, k)
        ids = torch.stack([bid.to(torch.int32), did.to(torch.int32)], 0).contiguous()
        return ops.rrf_fuse(ids, 10)
    hy_med, hy_best = timeit(hybrid, 10)
    ids = torch.stack([ops.bm25_topk(idx, qt, ql, k)[1].to(torch.int32), ops.dense_topk(x, q, k)[1].to(torch.int32)], 0).contiguous()
    rr_med, rr_best = timeit(lambda: ops.rrf_fuse(ids, 10), 20)
    rr_graph = graph_time(lambda: ops.rrf_fuse(ids, 10))
    hy_graph = graph_time(hybrid)
    emit(config="C2-rrf", batch=nq, lists=2, kl=k, top_k=10, ms_median=rr_med, ms_best=rr_best, ms_graph=rr_graph)
    emit(config="C2-hybrid", batch=nq, ms_median=hy_med, ms_best=hy_best, ms_graph=hy_graph, qps=nq / (hy_med * 1e-3),
         qps_graph=nq / (hy_graph * 1e-3),
         note="ms_median: eager calls from Python, one at a time; ms_graph: the same calls replayed from a CUDA graph")


def c2_plugin(dev):
    """C2 through the plugin classes, strings in / Document objects out: BM25Retriever +
    VectorStoreRetriever(B200VectorStore) + RRFusion behind MultiPathRetriever.invoke_batch
    (the reference's MultiPathRetriever takes one query per call, mutipath.py:37-93)."""
    from rag_arc_b200.core.retrieval.bm25 import BM25Retriever
    from rag_arc_b200.core.retrieval.dense import VectorStoreRetriever
    from rag_arc_b200.core.retrieval.mutipath import MultiPathRetriever
    from rag_arc_b200.core.utils.Fusion import RRFusion
    from rag_arc_b200.encapsulation.database.vector_db.VectorStore_B200 import B200VectorStore
    from rag_arc_b200.encapsulation.embeddings.pooled import TableEmbeddings
    n, d, nq, k = 100_000, 768, 256, 50
    toks, offs = synth.bm25_corpus_tokens(n)
    qtok = synth.bm25_queries_tokens(toks, offs, nq)
    texts = [" ".join(f"t{t}" for t in toks[offs[i]:offs[i + 1]]) + f" doc{i}" for i in range(n)]
    queries = [" ".join(f"t{t}" for t in row) for row in qtok]
    X = synth.dense_corpus_np(n, d)
    Q, _ = synth.dense_queries_np(X, nq)
    table = {q: Q[i] for i, q in enumerate(queries)}
    t0 = time.perf_counter()
    store = B200VectorStore.from_embeddings(texts, X, embedding=TableEmbeddings(table), metric="cosine",
                                            dtype="bfloat16", device=dev)
    bm = BM25Retriever.from_texts(texts, k=k, device=dev)
    hybrid = MultiPathRetriever([bm, VectorStoreRetriever(store, search_kwargs={"k": k})], RRFusion(),
                                top_k_per_retriever=k)
    build_s = time.perf_counter() - t0
    out = hybrid.invoke_batch(queries, top_k=10)           # first call builds the key tables
    assert len(out) == nq and all(len(o) == 10 for o in out)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        hybrid.invoke_batch(queries, top_k=10)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    for qi in range(3):                                   # warm: kernel variants and workspaces of nq = 1
        single = hybrid.invoke(queries[qi], top_k=10)
    t0 = time.perf_counter()
    for qi in range(20):
        hybrid.invoke(queries[qi], top_k=10)
    single_ms = (time.perf_counter() - t0) * 1e3 / 20
    single = hybrid.invoke(queries[0], top_k=10)
    emit(config="C2-plugin", docs=n, batch=nq, top_k=10, ms_median=ts[len(ts) // 2], ms_best=ts[0],
         qps=nq / (ts[len(ts) // 2] * 1e-3), single_query_ms=single_ms, build_s=build_s,
         same_first_doc=bool(single[0].content == out[0][0].content),
         note="wall clock, strings in -> Document objects out, MultiPathRetriever.invoke_batch")


def pool_case(dev):
    B, T, H = 256, 512, 768
    x = torch.randn((B, T, H), device=dev, dtype=torch.float16)
    lens = torch.randint(T // 4, T + 1, (B,), device=dev)
    mask = (torch.arange(T, device=dev)[None, :] < lens[:, None]).int()
    med, best = timeit(lambda: ops.pool_normalize(x, mask, "mean", True), 20)
    bytes_ = int(lens.sum().item()) * H * 2 + B * T * 4 + B * H * 4
    emit(config="pool-normalize", B=B, T=T, H=H, dtype="float16", ms_median=med, ms_best=best,
         algorithmic_bytes=bytes_, hbm_gbs=bytes_ / (med * 1e-3) / 1e9, frac_hbm_peak=bytes_ / (med * 1e-3) / 1e9 / PEAKS["hbm_gbs"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    want = set(args.only.split(",")) if args.only else None
    def on(tag):
        return want is None or tag in want
    if on("c1"):
        dense_case("C1", 10_000, 384, torch.float32, 100, 10, dev)
    if on("c2"):
        c2_hybrid(dev)
    if on("c2plugin"):
        c2_plugin(dev)
    if on("pool"):
        pool_case(dev)
    if on("c4"):
        dense_case("C4-shard(G=8)", 1_250_000, 1024, torch.float16, 1024, 100, dev)
    if on("c5"):
        for nq in (1, 8, 64):
            dense_case(f"C5-shard(G=8)-Q{nq}", 6_250_000, 768, torch.bfloat16, nq, 100, dev, iters=10)
    if on("x3"):
        x3_case(dev)
    if on("simt"):
        dense_case("fp32-100kx768-Q256", 100_000, 768, torch.float32, 256, 50, dev, iters=5)


if __name__ == "__main__":
    main()
