"""Generate ``tests/golden/*`` by EXECUTING THE REFERENCE'S OWN CODE (test infrastructure).

Run in the build container only (``/root/reference`` must exist):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden

What is live reference code here: ``core/utils/Fusion.py`` (RRFusion), ``core/retrieval/mutipath.py``
(MultiPathRetriever), ``core/retrieval/dense.py`` (VectorStoreRetriever),
``encapsulation/database/vector_db/VectorStore_Faiss.py`` + ``VectorStoreBase.py``
(FaissVectorStore incl. relevance scores, threshold filter, MMR), ``core/retrieval/bm25.py``
(BM25Retriever).  What is NOT: ``faiss`` and ``rank_bm25`` themselves (absent third-party
packages) - they are the numpy restatements in ``oracle/shims`` ("parity unpinned" for those,
see oracle/__init__.py).  The files written are small and committed; tests on the GPU box read
only them, never /root/reference.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


class TableEmbeddings:
    """Deterministic text -> vector lookup standing in for an embedding model."""

    def __init__(self, table):
        self.table = table

    def embed_documents(self, texts):
        return [self.table[t].tolist() for t in texts]

    def embed_query(self, text):
        return self.table[text].tolist()


def gen_dense_l2(ns):
    """Metric "l2" of the reference's FaissVectorStore (IndexFlatL2 behind it, squared distances
    ascending, _euclidean_relevance_score_fn) over the vectors of dense_small.npz."""
    import warnings
    z = np.load(os.path.join(GOLD, "dense_small.npz"))
    with open(os.path.join(GOLD, "dense_small.json")) as f:
        base = json.load(f)
    texts, queries = base["texts"], base["queries"]
    table = {t: v for t, v in zip(texts, z["vecs"])}
    table.update({t: v for t, v in zip(queries, z["qvecs"])})
    table["test"] = np.zeros(z["vecs"].shape[1], np.float32)
    emb = TableEmbeddings(table)
    out = {"source": "FaissVectorStore(metric='l2') executed live over oracle.dense.IndexFlatL2", "cases": []}
    for normalize in (False, True):
        store = ns.FaissVectorStore.from_texts(texts, emb, ids=[str(i) for i in range(len(texts))], metric="l2",
                                               normalize_L2=normalize)
        for k in (1, 4, 10):
            for qi, qt in enumerate(queries):
                res = store.similarity_search_with_score(qt, k)
                out["cases"].append({"kind": "similarity_with_score", "normalize_L2": normalize, "k": k, "query": qi,
                                     "ids": [int(doc.id) for doc, _ in res], "scores": [float(s) for _, s in res]})
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for qi, qt in enumerate(queries[:4]):
                rel = store.similarity_search_with_relevance_scores(qt, k=6)
                out["cases"].append({"kind": "relevance", "normalize_L2": normalize, "query": qi,
                                     "ids": [int(d_.id) for d_, _ in rel], "relevance": [float(s) for _, s in rel]})
        for qi, qt in enumerate(queries[:4]):
            docs = store.max_marginal_relevance_search(qt, k=3, fetch_k=10, lambda_mult=0.5)
            out["cases"].append({"kind": "mmr", "normalize_L2": normalize, "query": qi, "ids": [int(d_.id) for d_ in docs]})
    with open(os.path.join(GOLD, "dense_l2_small.json"), "w") as f:
        json.dump(out, f, indent=1)
    return len(out["cases"])


def gen_rrf(ns):
    rng = np.random.default_rng(20250101)
    cases = []

    def run(name, lists, top_k, k=60.0):
        res = [[ns.RetrievalResult(document=ns.Document(content=str(i), id=str(i)), score=1.0) for i in lst]
               for lst in lists]
        fused = ns.RRFusion(k=k).fuse(res, top_k)
        cases.append({"name": name, "lists": [list(map(int, l)) for l in lists], "top_k": top_k, "k": k,
                      "fused_ids": [int(r.document.content) for r in fused],
                      "fused_scores": [float(r.score) for r in fused],
                      "fused_ranks": [int(r.rank) for r in fused]})

    for t in range(12):
        L = int(rng.integers(1, 4))
        pool = int(rng.integers(20, 200))
        lists = [rng.permutation(pool)[:int(rng.integers(0, 51))].tolist() for _ in range(L)]
        run(f"random{t}", lists, int(rng.choice([1, 5, 10, 50])))
    # 1-ulp collisions between rank pairs: (6,39) vs (12,28) and (30,50) vs (39,39)
    a = list(range(100, 150)); b = list(range(200, 250))
    a[5] = b[38] = 1            # doc 1: ranks 6 and 39
    a[11] = b[27] = 2           # doc 2: ranks 12 and 28
    run("ulp_6_39_vs_12_28", [a, b], 50)
    run("ulp_6_39_vs_12_28_swapped", [b, a], 50)
    a = list(range(100, 150)); b = list(range(200, 250))
    a[29] = b[49] = 3           # doc 3: ranks 30 and 50
    a[38] = b[38] = 4           # doc 4: ranks 39 and 39
    run("ulp_30_50_vs_39_39", [a, b], 50)
    run("ulp_30_50_vs_39_39_swapped", [b, a], 50)
    run("all_ties_first_insertion_order", [[5, 6, 7], [8, 9, 10]], 6)
    run("identical_lists", [[1, 2, 3, 4], [1, 2, 3, 4]], 3)
    run("reverse_lists", [[1, 2, 3, 4], [4, 3, 2, 1]], 4)
    run("one_empty", [[], [3, 1, 2]], 10)
    run("k1", [[1, 2, 3], [3, 2, 1]], 3, k=1.0)
    run("k0p5", [[9, 8, 7, 6], [6, 7], [7]], 4, k=0.5)
    with open(os.path.join(GOLD, "rrf_reference.json"), "w") as f:
        json.dump({"source": "core/utils/Fusion.py RRFusion.fuse executed live", "cases": cases}, f, indent=1)
    return len(cases)


def gen_rrf_rows(ns):
    """The hybrid merge on Documents with DUPLICATED contents (inside one retriever's corpus and across
    retrievers): which Document object RRFusion.fuse hands back for a content - the last one seen while the
    lists are walked in order (Fusion.py:61).  Rows and per-retriever row -> content tables go in, (list, row)
    of every fused Document comes out: the vectors ragarc_rrf_fuse_rows and oracle.rrf.rrf_fuse_rows are
    checked against."""
    rng = np.random.default_rng(20250202)
    cases = []
    for t in range(16):
        L = int(rng.integers(1, 4))
        n_contents = int(rng.integers(8, 60))
        sizes = [int(rng.integers(10, 80)) for _ in range(L)]
        contents = [rng.integers(0, n_contents, n).tolist() for n in sizes]           # row -> content key
        rows = [rng.permutation(n)[:int(rng.integers(0, min(n, 50) + 1))].tolist() for n in sizes]
        top_k = int(rng.choice([1, 5, 10, 50]))
        res = [[ns.RetrievalResult(document=ns.Document(content=f"c{contents[l][r]}", id=f"{l}:{r}"), score=1.0)
                for r in rows[l]] for l in range(L)]
        fused = ns.RRFusion(k=60.0).fuse(res, top_k)
        cases.append({"name": f"dups{t}", "rows": rows, "contents": contents, "top_k": top_k, "k": 60.0,
                      "fused": [[int(x) for x in r.document.id.split(":")] for r in fused],
                      "fused_scores": [float(r.score) for r in fused]})
    with open(os.path.join(GOLD, "rrf_rows_reference.json"), "w") as f:
        json.dump({"source": "core/utils/Fusion.py RRFusion.fuse executed live on Documents with duplicated contents",
                   "cases": cases}, f, indent=1)
    return len(cases)


def gen_dense(ns):
    rng = np.random.default_rng(42)
    n, d = 300, 48
    texts = [f"doc {i} " + " ".join(f"w{int(w)}" for w in rng.integers(0, 50, 6)) for i in range(n)]
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    vecs[17] = vecs[3]                       # exact duplicate -> tie
    queries = [f"query {j}" for j in range(12)]
    qvecs = (vecs[rng.integers(0, n, 12)] + 0.3 * rng.standard_normal((12, d))).astype(np.float32)
    table = {t: v for t, v in zip(texts, vecs)}
    table.update({t: v for t, v in zip(queries, qvecs)})
    table["test"] = np.zeros(d, np.float32)
    emb = TableEmbeddings(table)
    out = {"texts": texts, "queries": queries, "cases": []}
    for metric in ("cosine", "ip"):
        store = ns.FaissVectorStore.from_texts(texts, emb, ids=[str(i) for i in range(n)], metric=metric)
        for k in (1, 4, 10):
            for qi, qt in enumerate(queries):
                res = store.similarity_search_with_score(qt, k)
                out["cases"].append({"kind": "similarity_with_score", "metric": metric, "k": k, "query": qi,
                                     "ids": [int(doc.id) for doc, _ in res],
                                     "scores": [float(s) for _, s in res]})
        r = ns.VectorStoreRetriever(vectorstore=store)      # default k=5 (dense.py:139)
        for qi, qt in enumerate(queries[:4]):
            out["cases"].append({"kind": "retriever_default", "metric": metric, "query": qi,
                                 "ids": [int(d_.id) for d_ in r.invoke(qt)]})
        if metric == "cosine":
            for thr in (0.0, 0.3):
                r = ns.VectorStoreRetriever(vectorstore=store, search_type="similarity_score_threshold",
                                            search_kwargs={"score_threshold": thr, "k": 8})
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    for qi, qt in enumerate(queries[:4]):
                        rel = store.similarity_search_with_relevance_scores(qt, k=8, score_threshold=thr)
                        out["cases"].append({"kind": "score_threshold", "metric": metric, "thr": thr, "query": qi,
                                             "ids": [int(d_.id) for d_ in r.invoke(qt)],
                                             "relevance": [float(s) for _, s in rel]})
            for qi, qt in enumerate(queries[:6]):
                docs = store.max_marginal_relevance_search(qt, k=4, fetch_k=12, lambda_mult=0.5)
                out["cases"].append({"kind": "mmr", "metric": metric, "query": qi, "k": 4, "fetch_k": 12,
                                     "lambda_mult": 0.5, "ids": [int(d_.id) for d_ in docs]})
    np.savez_compressed(os.path.join(GOLD, "dense_small.npz"), vecs=vecs, qvecs=qvecs)
    with open(os.path.join(GOLD, "dense_small.json"), "w") as f:
        json.dump(out, f)
    # a folder as the reference's own FaissVectorStore.save_local leaves it (:432-450): index.faiss
    # (flat layout, via the faiss shim) + index.pkl holding the REFERENCE's Document objects - the
    # input of rag_arc_b200.formats / B200VectorStore.load_local
    store = ns.FaissVectorStore.from_texts(texts[:60], emb, ids=[f"id{i}" for i in range(60)], metric="cosine",
                                           metadatas=[{"pos": i, "tags": ["a", i % 3]} for i in range(60)])
    store.save_local(os.path.join(GOLD, "ref_saved_store"), "index")
    return len(out["cases"])


def _hybrid_cases(ns, texts, queries, vecs, qvecs, bm):
    table = {}
    for t, v in zip(texts, vecs):
        table.setdefault(t, v)
    table.update({q: v for q, v in zip(queries, qvecs)})
    table["test"] = np.zeros(vecs.shape[1], np.float32)
    store = ns.FaissVectorStore.from_texts(texts, TableEmbeddings(table), ids=[str(i) for i in range(len(texts))])
    dense_r = ns.VectorStoreRetriever(vectorstore=store)

    class Failing(ns.BaseRetriever):
        def _get_relevant_documents(self, query, **kw):
            raise RuntimeError("boom")

    import contextlib
    import io
    out = []
    for name, retrievers in (("bm25+dense", [bm, dense_r]), ("dense+bm25", [dense_r, bm]),
                             ("bm25+fail+dense", [bm, Failing(), dense_r])):
        mp = ns.MultiPathRetriever(retrievers, top_k_per_retriever=50)
        for qi, q in enumerate(queries):
            for top_k in (10, 50):
                with contextlib.redirect_stdout(io.StringIO()):
                    docs = mp.invoke(q, top_k=top_k)
                out.append({"combo": name, "query": qi, "top_k": top_k,
                            "contents_idx": [texts.index(d_.content) for d_ in docs],
                            "ids": [int(d_.id) for d_ in docs]})
    return out


def _hybrid_tagged_cases(ns, texts, queries, vecs, qvecs):
    """Same corpus in both retrievers, but DIFFERENT Document objects per text (ids ``b<i>`` in the BM25
    retriever, ``d<i>`` in the vector store): the ids of the fused results say which retriever's Document the
    reference hands back - the last one seen in the walk over the lists (Fusion.py:61)."""
    from oracle.ref_loader import make_bm25_retriever
    bm = make_bm25_retriever(ns, texts, k=5)
    for doc in bm.docs:
        doc.id = "b" + doc.id
    table = {}
    for t, v in zip(texts, vecs):
        table.setdefault(t, v)
    table.update({q: v for q, v in zip(queries, qvecs)})
    store = ns.FaissVectorStore.from_texts(texts, TableEmbeddings(table), ids=[f"d{i}" for i in range(len(texts))])
    dense_r = ns.VectorStoreRetriever(vectorstore=store)
    out = []
    for name, retrievers in (("bm25+dense", [bm, dense_r]), ("dense+bm25", [dense_r, bm])):
        mp = ns.MultiPathRetriever(retrievers, top_k_per_retriever=50)
        for qi, q in enumerate(queries):
            for top_k in (10, 50):
                out.append({"combo": name, "query": qi, "top_k": top_k, "ids": [d_.id for d_ in mp.invoke(q, top_k=top_k)]})
    return out


def gen_bm25_hybrid(ns):
    from oracle.ref_loader import make_bm25_retriever
    rng = np.random.default_rng(7)
    vocab = [f"w{i}" for i in range(60)]
    p = 1.0 / np.arange(1, 61) ** 1.1
    p /= p.sum()
    texts = [" ".join(rng.choice(vocab, size=int(rng.integers(3, 30)), p=p)) for _ in range(250)]
    texts[40] = texts[12]                    # duplicate content: RRF dedups by content (Fusion.py:59)
    queries = ["w0 w3 w7", "w1 w1 w59", "w5 w40 w41 w42", "nosuchtoken w2", "w58", "w0"]
    bm = make_bm25_retriever(ns, texts, k=5)
    # the file the reference's own BM25Retriever.save_to_disk writes (bm25.py:550-576): input of
    # rag_arc_b200.formats.load_reference_bm25_state / BM25Retriever.load_from_disk
    os.makedirs(os.path.join(GOLD, "ref_saved_bm25"), exist_ok=True)
    bm.save_to_disk(os.path.join(GOLD, "ref_saved_bm25"))
    out = {"texts": texts, "queries": queries, "bm25": [], "hybrid": []}
    for qi, q in enumerate(queries):
        scores = bm.get_scores(q)
        for k in (3, 10, 50):
            docs = bm.invoke(q, k=k)
            out["bm25"].append({"query": qi, "k": k, "ids": [int(d.id) for d in docs]})
        out["bm25"].append({"query": qi, "scores": [float(s) for s in scores]})
    # hybrid: BM25 + dense (table embeddings) through the reference MultiPathRetriever
    d = 32
    vecs = rng.standard_normal((len(texts), d)).astype(np.float32)
    vecs[40] = vecs[12]
    qvecs = rng.standard_normal((len(queries), d)).astype(np.float32)
    out["hybrid"] = _hybrid_cases(ns, texts, queries, vecs, qvecs, bm)
    np.savez_compressed(os.path.join(GOLD, "hybrid_small.npz"), vecs=vecs, qvecs=qvecs)
    with open(os.path.join(GOLD, "bm25_hybrid_small.json"), "w") as f:
        json.dump(out, f)

    # second set built to be free of BM25 ties inside the top 50 (every document has a distinct
    # length and contains the frequent words), so the reference's unstable argsort order is
    # irrelevant and its fused ranking can be compared exactly
    vocab2 = [f"v{i}" for i in range(25)]
    p2 = 1.0 / np.arange(1, 26) ** 0.8
    p2 /= p2.sum()
    texts2 = [" ".join(rng.choice(vocab2, size=30 + i, p=p2)) for i in range(300)]
    queries2 = ["v0 v1", "v2", "v0 v3 v5 v0", "v1 v4"]
    bm2 = make_bm25_retriever(ns, texts2, k=5)
    vecs2 = rng.standard_normal((len(texts2), d)).astype(np.float32)
    qvecs2 = rng.standard_normal((len(queries2), d)).astype(np.float32)
    out2 = {"texts": texts2, "queries": queries2, "bm25": [], "hybrid": _hybrid_cases(ns, texts2, queries2, vecs2, qvecs2, bm2)}
    for qi, q in enumerate(queries2):
        sc = np.asarray(bm2.get_scores(q))
        top = np.sort(sc)[::-1][:52]
        assert len(np.unique(top)) == len(top), "tie inside the top 52 - regenerate with another seed"
        out2["bm25"].append({"query": qi, "scores": [float(x) for x in sc]})
        out2["bm25"].append({"query": qi, "k": 50, "ids": [int(d_.id) for d_ in bm2.invoke(q, k=50)]})
    out2["hybrid_tagged"] = _hybrid_tagged_cases(ns, texts2, queries2, vecs2, qvecs2)
    np.savez_compressed(os.path.join(GOLD, "hybrid_tiefree.npz"), vecs=vecs2, qvecs=qvecs2)
    with open(os.path.join(GOLD, "hybrid_tiefree.json"), "w") as f:
        json.dump(out2, f)
    return len(out["bm25"]) + len(out["hybrid"]) + len(out2["hybrid"])


def gen_adjacent_cosine(ns):
    """Runs the reference's own calculate_cosine_distances (spliter.py:354-372) on seeded sentence
    embeddings (including a zero row and a duplicated row) and stores inputs + outputs."""
    from core.file_management.chunker.spliter import calculate_cosine_distances
    rng = np.random.default_rng(77)
    emb = rng.standard_normal((40, 96)).astype(np.float32)
    emb[7] = 0.0                      # zero row: similarity nan -> 0 -> distance 1
    emb[20] = emb[19]                 # identical neighbours: distance ~0
    emb[30:] *= 1e-3
    sentences = [{"sentence": str(i), "combined_sentence_embedding": e.tolist()} for i, e in enumerate(emb)]
    distances, sentences = calculate_cosine_distances(sentences)
    assert all(sentences[i]["distance_to_next"] == distances[i] for i in range(len(distances)))
    np.savez_compressed(os.path.join(GOLD, "adjacent_cosine.npz"), emb=emb,
                        distances=np.asarray(distances, dtype=np.float64))
    return len(distances)


def main():
    from oracle import ref_loader
    ns = ref_loader.load()
    os.makedirs(GOLD, exist_ok=True)
    print("rrf cases:", gen_rrf(ns))
    print("rrf rows cases:", gen_rrf_rows(ns))
    print("dense cases:", gen_dense(ns))
    print("bm25+hybrid cases:", gen_bm25_hybrid(ns))
    print("adjacent cosine pairs:", gen_adjacent_cosine(ns))
    print("dense l2 cases:", gen_dense_l2(ns))


if __name__ == "__main__":
    main()
