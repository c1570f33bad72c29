"""cProfile of MultiPathRetriever.invoke_batch at the C2 shape (100k docs, 256 queries, top-50 each, RRF top-10)."""
import cProfile, io, os, pstats, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rag_arc_b200 import synth
from rag_arc_b200.core.retrieval.bm25 import BM25Retriever
from rag_arc_b200.core.retrieval.dense import VectorStoreRetriever
from rag_arc_b200.core.retrieval.mutipath import MultiPathRetriever
from rag_arc_b200.core.utils.Fusion import RRFusion
from rag_arc_b200.encapsulation.database.vector_db.VectorStore_B200 import B200VectorStore
from rag_arc_b200.encapsulation.embeddings.pooled import TableEmbeddings
dev = torch.device("cuda:0")
n, d, nq, k = 100_000, 768, 256, 50
toks, offs = synth.bm25_corpus_tokens(n)
qtok = synth.bm25_queries_tokens(toks, offs, nq)
texts = [t + f" doc{i}" for i, t in enumerate(synth.tokens_to_texts(toks, offs))]
queries = [" ".join(f"t{t}" for t in row) for row in qtok]
X = synth.dense_corpus_np(n, d); Q, _ = synth.dense_queries_np(X, nq)
store = B200VectorStore.from_embeddings(texts, X, embedding=TableEmbeddings({s: Q[i] for i, s in enumerate(queries)}),
                                        metric="cosine", dtype="bfloat16", device=dev)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    bm = BM25Retriever.from_texts(texts, k=k, device=dev)
hyb = MultiPathRetriever([bm, VectorStoreRetriever(store, search_kwargs={"k": k})], RRFusion(), top_k_per_retriever=k)
for _ in range(3):
    hyb.invoke_batch(queries, top_k=10)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); hyb.invoke_batch(queries, top_k=10); ts.append((time.perf_counter() - t0) * 1e3)
print("invoke_batch ms: median %.3f best %.3f" % (sorted(ts)[10], min(ts)))
# phases of one call: host work until the fused merge is enqueued / waiting for the device + the one
# device->host transfer / building the result lists; and the same order reversed (dense first)
from rag_arc_b200 import ops as _ops
marks = {}
_fuse, _unpack = hyb.fusion_method.fuse_rows_batch, _ops.unpack_fused_rows
def fuse_marked(*a, **kw):
    r = _fuse(*a, **kw); marks["enqueued"] = time.perf_counter(); return r
def unpack_marked(*a, **kw):
    marks["synced"] = time.perf_counter(); return _unpack(*a, **kw)
hyb.fusion_method.fuse_rows_batch = fuse_marked; _ops.unpack_fused_rows = unpack_marked
ph = []
for _ in range(20):
    t0 = time.perf_counter(); hyb.invoke_batch(queries, top_k=10); t1 = time.perf_counter()
    ph.append(((marks["enqueued"] - t0) * 1e3, (marks["synced"] - marks["enqueued"]) * 1e3, (t1 - marks["synced"]) * 1e3))
ph = np.median(np.array(ph), axis=0)
print("phases ms (median): host work until everything is enqueued %.3f | wait for the device + D2H %.3f | result lists %.3f" % tuple(ph))
hyb.fusion_method.fuse_rows_batch = _fuse; _ops.unpack_fused_rows = _unpack
rev = MultiPathRetriever([hyb.retrievers[1], hyb.retrievers[0]], RRFusion(), top_k_per_retriever=k)
for _ in range(3):
    rev.invoke_batch(queries, top_k=10)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); rev.invoke_batch(queries, top_k=10); ts.append((time.perf_counter() - t0) * 1e3)
print("dense first, then BM25: invoke_batch ms: median %.3f best %.3f" % (sorted(ts)[10], min(ts)))
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    hyb.invoke_batch(queries, top_k=10)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
