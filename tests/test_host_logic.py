"""CPU: host-side logic of the plugins (no kernels run): BM25 index build against the oracle,
retriever argument validation and orchestration, store bookkeeping errors."""
import contextlib
import io
import warnings

import numpy as np
import pytest

from oracle import bm25 as obm25
from rag_arc_b200 import synth
from rag_arc_b200.core.retrieval.base import BaseRetriever
from rag_arc_b200.core.retrieval.bm25 import BM25Retriever, default_preprocessing_func
from rag_arc_b200.core.retrieval.bm25_index import Bm25Index
from rag_arc_b200.core.retrieval.dense import VectorStoreRetriever
from rag_arc_b200.core.retrieval.mutipath import MultiPathRetriever
from rag_arc_b200.core.utils.data_model import Document
from rag_arc_b200.core.utils.Fusion import FusionMethod, RetrievalResult
from rag_arc_b200.encapsulation.database.vector_db.VectorStore_B200 import B200VectorStore
from rag_arc_b200.encapsulation.embeddings.pooled import HashEmbeddings
from rag_arc_b200.sharded import shard_bounds


def test_bm25_index_tables_bit_identical_to_oracle():
    rng = np.random.default_rng(2)
    vocab = [f"w{i}" for i in range(120)]
    p = 1.0 / np.arange(1, 121) ** 1.2; p /= p.sum()
    corpus = [rng.choice(vocab, size=int(rng.integers(1, 50)), p=p).tolist() for _ in range(600)]
    ref = obm25.BM25Okapi(corpus)
    csr = obm25.Bm25Csr(corpus)
    idx = Bm25Index.from_token_lists(corpus, device="cpu")
    assert idx.avgdl == ref.avgdl and idx.average_idf == ref.average_idf and idx.n_docs == 600
    for tok, ti in idx.vocab.items():
        assert idx.idf_np[ti] == ref.idf[tok]
    assert np.array_equal(idx.doc_norm_np.view(np.uint64), csr.doc_norm.view(np.uint64))
    assert np.array_equal(idx.indptr_np, csr.indptr) and np.array_equal(idx.post_doc_np, csr.post_doc)
    assert np.array_equal(idx.post_tf_np, csr.post_tf)
    qt, ql = idx.encode_queries([["w0", "nope", "w3", "w0"], []])
    assert qt.shape == (2, 4) and ql.tolist() == [4, 0]
    assert qt[0].tolist() == [idx.vocab["w0"], -1, idx.vocab["w3"], idx.vocab["w0"]]


def test_bm25_index_from_token_ids_matches_from_token_lists():
    toks, offs = synth.bm25_corpus_tokens(300, vocab=500, seed=5)
    a = Bm25Index.from_token_ids(toks, offs, device="cpu")
    b = Bm25Index.from_token_lists([[f"t{t}" for t in toks[offs[i]:offs[i + 1]]] for i in range(300)], device="cpu")
    assert np.array_equal(a.indptr_np, b.indptr_np) and np.array_equal(a.post_doc_np, b.post_doc_np)
    assert np.array_equal(a.idf_np.view(np.uint64), b.idf_np.view(np.uint64))
    q = synth.bm25_queries_tokens(toks, offs, 4, 8, seed=6)
    qa, _ = a.encode_query_ids(q)
    qb, _ = b.encode_queries([[f"t{t}" for t in row] for row in q])
    assert qa.tolist() == qb.tolist()


def test_bm25_retriever_constructor_validation():
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        BM25Retriever()
        assert any("whitespace" in str(x.message) for x in w)
    with pytest.raises(ValueError):
        BM25Retriever(k=0, warn_default_preprocess=False)
    with pytest.raises(ValueError):
        BM25Retriever(preprocess_func=3, warn_default_preprocess=False)
    r = BM25Retriever(warn_default_preprocess=False)
    assert r.k == 5 and r.preprocess_func is default_preprocessing_func
    with pytest.raises(ValueError):
        r.invoke("q")                               # vectorizer not initialised
    with pytest.raises(ValueError):
        BM25Retriever.from_texts([])
    with pytest.raises(ValueError):
        BM25Retriever.from_texts(["a"], metadatas=[{}, {}])
    with pytest.raises(ValueError):
        r.update_k(0)
    assert r.get_name() == "BM25Retriever" and r.get_document_count() == 0


def test_vector_store_retriever_validation():
    store = B200VectorStore(embedding=HashEmbeddings(16))
    with pytest.raises(ValueError):
        VectorStoreRetriever(vectorstore=store, search_type="nope")
    with pytest.raises(ValueError):
        VectorStoreRetriever(vectorstore=store, search_type="similarity_score_threshold")
    with pytest.raises(ValueError):
        VectorStoreRetriever(vectorstore=store, search_type="similarity_score_threshold",
                             search_kwargs={"score_threshold": 1.5})
    r = VectorStoreRetriever(vectorstore=store, search_kwargs={"k": 3})
    assert r.invoke("anything") == []               # empty store -> []
    assert r.get_name() == "B200VectorStoreRetriever"
    assert r.get_vectorstore_info()["embedding_class"] == "HashEmbeddings"
    with pytest.raises(ValueError):
        r.update_search_params(search_type="bad")


def test_vector_store_argument_errors_and_relevance_fns():
    with pytest.raises(ValueError):
        B200VectorStore(embedding=None, dtype="int8")
    s = B200VectorStore(embedding=HashEmbeddings(8), metric="nope")
    with pytest.raises(ValueError):
        s._create_index(8)
    with pytest.raises(ValueError):
        B200VectorStore(embedding=None, index_type="ivf")._create_index(8)
    assert B200VectorStore(embedding=None, metric="cosine")._select_relevance_score_fn()(0.25) == 0.75
    ip = B200VectorStore(embedding=None, metric="ip")._select_relevance_score_fn()
    assert ip(0.25) == 0.75 and ip(-2.0) == 2.0
    assert s.delete([]) is True and s.similarity_search("q") == [] and s.get_by_ids(["x"]) == []


class _Stub(BaseRetriever):
    def __init__(self, docs, fail=False):
        super().__init__()
        self.docs, self.fail, self.seen = docs, fail, None

    def _get_relevant_documents(self, query, **kw):
        self.seen = kw
        if self.fail:
            raise RuntimeError("boom")
        return self.docs[:kw["k"]]


class _ConcatFusion(FusionMethod):
    def fuse(self, results, top_k):
        flat = [r for lst in results for r in lst]
        return flat[:top_k]


def test_multipath_orchestration_matches_reference_semantics():
    a = _Stub([Document(content=f"a{i}") for i in range(80)])
    b = _Stub([], fail=True)
    c = _Stub([Document(content=f"c{i}") for i in range(3)])
    mp = MultiPathRetriever([a, b, c], fusion_method=_ConcatFusion(), top_k_per_retriever=50)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        docs = mp.invoke("q", top_k=60, extra=1)
    assert "_Stub" in buf.getvalue() and "boom" in buf.getvalue()       # failure is printed, not raised
    assert a.seen == {"top_k": 60, "extra": 1, "k": 50}                  # k forced, other kwargs forwarded
    assert [d.content for d in docs] == [f"a{i}" for i in range(50)] + ["c0", "c1", "c2"]
    assert len(mp.invoke("q")) == 10                                     # top_k defaults to 10
    with contextlib.redirect_stdout(io.StringIO()):
        assert MultiPathRetriever([b], fusion_method=_ConcatFusion()).invoke("q") == []
    mp.remove_retriever("_Stub"); assert len(mp.retrievers) == 2


def test_shard_bounds_cover_rows_exactly_once():
    for n, g in [(1_000_000, 8), (10, 3), (7, 8), (0, 2), (1025, 4)]:
        spans = [shard_bounds(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo <= hi for lo, hi in spans)


def test_bm25_doc_range_shards_partition_the_postings_and_keep_global_statistics():
    """Bm25Index.shard (host side of the multi-GPU BM25): the shards' postings are a partition of the
    full index's, local doc ids are offset by id_base, idf / vocabulary are the full corpus's."""
    import numpy as np
    from rag_arc_b200.core.retrieval.bm25_index import Bm25Index
    rng = np.random.default_rng(4)
    words = [f"w{i}" for i in range(30)]
    corpus = [[words[j] for j in rng.integers(0, 30, rng.integers(0, 9))] for _ in range(57)]
    full = Bm25Index.from_token_lists(corpus, device=None)
    assert full.device is None
    bounds = [(0, 20), (20, 21), (21, 57)]
    shards = [full.shard(lo, hi, "cpu") for lo, hi in bounds]
    V = len(full.indptr_np) - 1
    for t in range(V):
        docs, tfs, vals = [], [], []
        for sh in shards:
            a, b = sh.indptr_np[t], sh.indptr_np[t + 1]
            docs += (sh.post_doc_np[a:b].astype(np.int64) + sh.id_base).tolist()
            tfs += sh.post_tf_np[a:b].tolist(); vals += sh.post_val_np[a:b].tolist()
        a, b = full.indptr_np[t], full.indptr_np[t + 1]
        assert docs == full.post_doc_np[a:b].tolist() and tfs == full.post_tf_np[a:b].tolist()
        assert vals == full.post_val_np[a:b].tolist()
    for sh, (lo, hi) in zip(shards, bounds):
        assert sh.n_docs == hi - lo and sh.id_base == lo
        assert sh.idf_np is full.idf_np and sh.vocab is full.vocab and sh.avgdl == full.avgdl
        assert np.array_equal(sh.doc_norm_np, full.doc_norm_np[lo:hi])
        assert sh.indptr.dtype == __import__("torch").int64 and sh.post_val.shape[0] == sh.post_doc.shape[0]


def test_async_wrappers_and_batch_defaults_of_the_plugin_bases():
    """base.py:52-67,82-96 / embeddings/base.py:36-61: the async forms run the synchronous methods in
    a thread pool; invoke_batch / embed_documents_array defaults fall back to the single-item API."""
    import asyncio
    import numpy as np
    from rag_arc_b200.core.file_management.embeddings.base import Embeddings
    from rag_arc_b200.core.retrieval.base import BaseRetriever
    from rag_arc_b200.core.utils.data_model import Document

    class E(Embeddings):
        def embed_documents(self, texts):
            return [[float(len(t)), 1.0] for t in texts]

        def embed_query(self, text):
            return [float(len(text)), 2.0]

    class R(BaseRetriever):
        def _get_relevant_documents(self, query, **kwargs):
            return [Document(content=query.upper(), metadata={"k": kwargs.get("k")})]

    e, r = E(anything="ignored"), R(search_kwargs={"k": 3}, tags=["t"])
    assert asyncio.run(e.aembed_documents(["ab", "c"])) == [[2.0, 1.0], [1.0, 1.0]]
    assert asyncio.run(e.aembed_query("abc")) == [3.0, 2.0]
    arr = e.embed_documents_array(["ab", "c"])
    assert arr.dtype == np.float32 and arr.tolist() == [[2.0, 1.0], [1.0, 1.0]]
    assert r.search_kwargs == {"k": 3} and r.tags == ["t"] and r.get_name() == "R"
    assert asyncio.run(r.ainvoke("x", k=2))[0].metadata == {"k": 2}
    assert [d[0].content for d in r.invoke_batch(["a", "b"])] == ["A", "B"]
    d = Document("text", {"a": 1}, "id1")                      # positional construction, as the reference allows
    assert Document.from_dict(d.to_dict()) == d and Document("x").metadata == {} and Document("x").id is None
    with __import__("pytest").raises(TypeError):
        BaseRetriever()                                        # abstract


def test_native_query_encoding_equals_python_split_and_lookup():
    """ragarc_vocab_encode_split[0]: the batch tokenise + lookup + pack behind the C ABI must give exactly
    what the reference's per-query Python does - ``query.split()`` (every str.isspace() character, UTF-8
    multi-byte ones included) and a dictionary lookup per token (-1 = out of vocabulary) - for empty
    queries, long queries (staging row regrown), repeated tokens, texts containing NUL, and random text."""
    import random
    import torch
    from rag_arc_b200.core.retrieval.bm25_index import Bm25Index
    spaces = [" ", "\t", "\n", "\x0b", "\x0c", "\r", "\x1c", "\x1d", "\x1e", "\x1f", "\x85", "\xa0", "\u1680", "\u2000",
              "\u2005", "\u200a", "\u2028", "\u2029", "\u202f", "\u205f", "\u3000"]
    assert all(c.isspace() for c in spaces)
    non_spaces = ["\u200b", "\u180e", "\ufeff", "\x00", "\x7f", "\xad"]          # look-alikes that str.split() keeps
    assert not any(c.isspace() for c in non_spaces)
    words = ["alpha", "beta", "gamma", "\u00fcn\u00ef", "\u65e5\u672c\u8a9e", "x", "a\u200bb", "\ufeffbom", "t1", "t22"]
    idx = Bm25Index.from_token_lists([words[:5], words[5:]], device=None)
    idx.device = torch.device("cpu")
    rnd = random.Random(7)
    texts = ["", " ", "alpha", "alpha  beta\tgamma zzz", "x " * 70, "a\u200bb \ufeffbom", "alpha\x00beta x"]
    for _ in range(300):
        parts = []
        for _ in range(rnd.randint(0, 12)):
            parts.append(rnd.choice(words + ["oov", "\u00e9"]))
            parts.append("".join(rnd.choice(spaces + non_spaces[:2]) for _ in range(rnd.randint(1, 3))))
        texts.append("".join(parts))
    for batch in (texts, texts[:6], texts[6:7], []):
        qt, ql = idx.encode_texts(batch)
        want = [[idx.vocab.get(t, -1) for t in s.split()] for s in batch]
        assert ql.tolist() == [len(w) for w in want]
        for i, row in enumerate(want):
            assert qt[i, :len(row)].tolist() == row, (batch[i], qt[i].tolist(), row)
            assert bool((qt[i, len(row):] == -1).all())
        if batch:
            qt2, ql2 = idx.encode_queries([s.split() for s in batch])
            w = min(qt.shape[1], qt2.shape[1])
            assert torch.equal(ql, ql2) and torch.equal(qt[:, :w], qt2[:, :w])
