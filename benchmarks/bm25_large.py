"""BM25 over a corpus whose postings do not fit L2 (2M documents, ~1 GB of postings): the regime
where the scatter kernel is really HBM-bound.  One JSON line: ms per 256-query batch (CUDA-graph
replay), algorithmic posting bytes per batch and the fraction of the measured HBM peak."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rag_arc_b200 import ops, synth
from rag_arc_b200.core.retrieval.bm25_index import Bm25Index
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
dev = torch.device("cuda:0")
n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nq, k = 256, 50
t0 = time.perf_counter()
toks, offs = synth.bm25_corpus_tokens(n_docs, vocab=200_000, mean_len=60.0)
qtok = synth.bm25_queries_tokens(toks, offs, nq)
idx = Bm25Index.from_token_ids(toks, offs, device=dev)
build_s = time.perf_counter() - t0
qt, ql = idx.encode_query_ids(qtok)
df = np.diff(idx.indptr_np)
sum_df = int(sum(df[t] for row in qt.cpu().numpy() for t in row if t >= 0))
fn = lambda: ops.bm25_topk(idx, qt, ql, k)
side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
scope = ops.WorkspaceScope()
with scope, torch.cuda.stream(side):
    fn(); fn(); side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        out = fn()
torch.cuda.current_stream(dev).wait_stream(side)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
# spot parity: two queries against the host restatement over the same CSR arrays
sc, ids = out
for qi in (0, nq - 1):
    s = np.zeros(n_docs)
    for t in qt[qi].cpu().numpy():
        if t >= 0:
            a, b = idx.indptr_np[t], idx.indptr_np[t + 1]
            docs = idx.post_doc_np[a:b]; tf = idx.post_tf_np[a:b].astype(np.int64)
            s[docs] += idx.idf_np[t] * (tf * (idx.k1 + 1) / (tf + idx.doc_norm_np[docs]))
    top = np.lexsort((np.arange(n_docs), -s))[:k]
    assert ids[qi].cpu().tolist() == top.tolist() and np.array_equal(sc[qi].cpu().numpy().view(np.uint64), s[top].view(np.uint64)), qi
bytes_alg = sum_df * 12
print(json.dumps({"config": "bm25-large", "docs": n_docs, "nnz": int(df.sum()), "postings_bytes": int(df.sum()) * 12, "batch": nq, "k": k,
                  "ms_per_batch": ms, "qps": nq / (ms * 1e-3), "postings_per_query": sum_df / nq,
                  "algorithmic_bytes_per_batch": bytes_alg, "achieved_gbs": bytes_alg / (ms * 1e-3) / 1e9,
                  "hbm_peak_gbs": peaks["hbm_gbs"], "frac_hbm_peak": bytes_alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                  "parity": "2 queries bit-exact against the host CSR restatement", "index_build_s": build_s}))
