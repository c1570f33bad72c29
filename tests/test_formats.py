"""On-disk formats of an existing RAG-ARC deployment (rag_arc_b200/formats.py): the FAISS flat-index
layout - pinned by a hand-assembled byte string that follows faiss/impl/index_write.cpp field by
field - and the pickle sidecar written by the REFERENCE's own FaissVectorStore.save_local
(tests/golden/ref_saved_store, produced by oracle/gen_golden.py running the reference's code)."""
import os
import pickle
import struct

import numpy as np
import pytest

from rag_arc_b200 import formats
from rag_arc_b200.core.utils.data_model import Document

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hand_built(rows, fourcc, metric_type, metric_arg=None):
    n, d = rows.shape
    b = fourcc
    b += struct.pack("<i", d)                       # Index::d (int)
    b += struct.pack("<q", n)                       # Index::ntotal (idx_t = int64)
    b += struct.pack("<q", 1 << 20) * 2             # two dummy idx_t
    b += struct.pack("<?", True)                    # is_trained (bool, one byte)
    b += struct.pack("<i", metric_type)             # MetricType (int)
    if metric_arg is not None:
        b += struct.pack("<f", metric_arg)
    b += struct.pack("<Q", n * d)                   # WRITEXBVECTOR: number of 4-byte units
    b += rows.astype("<f4").tobytes()
    return b


def test_faiss_flat_layout_hand_assembled(tmp_path):
    rows = np.arange(15, dtype=np.float32).reshape(5, 3) / 7
    for fourcc, mt, want in ((b"IxFI", 0, "ip"), (b"IxF2", 1, "l2")):
        p = tmp_path / f"{want}.faiss"
        p.write_bytes(_hand_built(rows, fourcc, mt))
        got, metric = formats.read_faiss_flat(str(p))
        assert metric == want and got.dtype == np.float32 and np.array_equal(got, rows)
        q = tmp_path / f"{want}_out.faiss"
        formats.write_faiss_flat(str(q), rows, want)
        assert q.read_bytes() == p.read_bytes()          # writer emits exactly the same bytes
    # empty index, and the generic fourcc that carries the metric in the header
    p = tmp_path / "empty.faiss"
    p.write_bytes(_hand_built(np.zeros((0, 8), np.float32), b"IxFl", 0))
    got, metric = formats.read_faiss_flat(str(p))
    assert got.shape == (0, 8) and metric == "ip"


@pytest.mark.parametrize("blob,msg", [
    (b"IwFl" + b"\0" * 64, "not a FAISS flat index"),                 # an IVF file
    (b"IxFI" + b"\0" * 10, "truncated index header"),
    (_hand_built(np.ones((2, 2), np.float32), b"IxFI", 0)[:-4], "truncated vector block"),
    (_hand_built(np.ones((2, 2), np.float32), b"IxFl", 4, 3.0), "neither inner product nor L2"),
])
def test_faiss_flat_reader_rejects_what_it_does_not_understand(tmp_path, blob, msg):
    p = tmp_path / "x.faiss"
    p.write_bytes(blob)
    with pytest.raises(ValueError, match=msg):
        formats.read_faiss_flat(str(p))


def test_reads_a_folder_saved_by_the_reference_itself():
    """index.faiss + index.pkl as FaissVectorStore.save_local left them (reference Document class
    inside the pickle): rows equal the normalised embeddings the reference indexed, Documents come
    back as this package's Document with content / metadata / id intact."""
    folder = os.path.join(GOLD, "ref_saved_store")
    rows, metric = formats.read_faiss_flat(os.path.join(folder, "index.faiss"))
    side = formats.load_reference_sidecar(os.path.join(folder, "index.pkl"))
    z = np.load(os.path.join(GOLD, "dense_small.npz"))
    want = z["vecs"][:60].copy()
    want *= (1.0 / np.sqrt((want * want).sum(1, keepdims=True))).astype(np.float32)
    assert metric == "ip" and rows.shape == (60, 48) and np.allclose(rows, want, rtol=0, atol=1e-6)
    assert side["metric"] == "cosine" and side["index_type"] == "flat" and side["normalize_L2"] in (False, True)
    assert side["index_to_docstore_id"] == {i: f"id{i}" for i in range(60)}
    doc = side["docstore"]["id7"]
    assert type(doc) is Document and doc.id == "id7" and doc.metadata == {"pos": 7, "tags": ["a", 1]}
    assert doc.content.startswith("doc 7 ")


def test_sidecar_loader_refuses_foreign_globals(tmp_path):
    """The reference loads its sidecar with pickle.load (VectorStore_Faiss.py:462), which executes
    whatever the file names; ours only ever constructs Document objects."""
    p = tmp_path / "evil.pkl"
    p.write_bytes(pickle.dumps({"docstore": {}, "index_to_docstore_id": {}, "fn": os.getcwd}))
    with pytest.raises(pickle.UnpicklingError):
        formats.load_reference_sidecar(str(p))
    ok = tmp_path / "ok.pkl"
    ok.write_bytes(pickle.dumps({"docstore": {"a": Document("x", {"s": {1, 2}}, "a")}, "index_to_docstore_id": {0: "a"},
                                 "index_type": "flat", "metric": "ip", "normalize_L2": False}))
    side = formats.load_reference_sidecar(str(ok))
    assert side["docstore"]["a"].metadata == {"s": {1, 2}}
    bad = tmp_path / "bad.pkl"
    bad.write_bytes(pickle.dumps([1, 2, 3]))
    with pytest.raises(ValueError):
        formats.load_reference_sidecar(str(bad))


def test_loads_the_bm25_state_saved_by_the_reference_and_rebuilds_identical_statistics():
    """tests/golden/ref_saved_bm25/bm25.pkl was dill-dumped by the reference's own
    BM25Retriever.save_to_disk (its vectorizer is a rank_bm25.BM25Okapi, its docs the reference's
    Documents).  Loading it here - without the reference or rank_bm25 importable - must rebuild an
    index whose idf table and average length equal the pickled ones bit for bit."""
    import json
    import sys
    from rag_arc_b200.core.retrieval.bm25 import BM25Retriever
    path = os.path.join(GOLD, "ref_saved_bm25", "bm25.pkl")
    assert "rank_bm25" not in sys.modules or "oracle" in (getattr(sys.modules["rank_bm25"], "__file__", "") or "")
    st = formats.load_reference_bm25_state(path)
    vec = st["vectorizer"]
    assert isinstance(vec, formats.ForeignBM25) and st["k"] == 5 and all(type(d) is Document for d in st["docs"])
    texts = json.load(open(os.path.join(GOLD, "bm25_hybrid_small.json")))["texts"]
    assert [d.content for d in st["docs"]] == texts and [d.id for d in st["docs"]] == [str(i) for i in range(len(texts))]
    r = BM25Retriever.load_from_disk(path, device="cpu")
    ix = r.vectorizer.index
    assert r.k == 5 and len(r.docs) == len(texts) and ix.n_docs == vec.corpus_size
    assert ix.avgdl == vec.avgdl and (ix.k1, ix.b, ix.epsilon) == (vec.k1, vec.b, vec.epsilon)
    ours = {tok: ix.idf_np[ti] for tok, ti in ix.vocab.items()}
    assert ours.keys() == vec.idf.keys()
    assert all(np.float64(ours[t]).tobytes() == np.float64(vec.idf[t]).tobytes() for t in ours)
    assert ix.doc_len_np.tolist() == list(vec.doc_len)
    with pytest.raises(IOError):
        BM25Retriever.load_from_disk(os.path.join(GOLD, "nope.pkl"))


def test_own_bm25_state_round_trips_through_the_same_loader(tmp_path):
    from rag_arc_b200.core.retrieval.bm25 import BM25Retriever
    texts = ["alpha beta gamma", "beta beta delta", "gamma epsilon", "zeta"]
    r = BM25Retriever.from_texts(texts, ids=["a", "b", "c", "d"], k=2, bm25_params={"k1": 1.2}, device="cpu")
    r.save_to_disk(str(tmp_path))
    back = BM25Retriever.load_from_disk(str(tmp_path / "bm25.pkl"), device="cpu")
    assert back.k == 2 and [d.id for d in back.docs] == ["a", "b", "c", "d"] and back.bm25_params == {"k1": 1.2}
    a, b = r.vectorizer.index, back.vectorizer.index
    assert a.k1 == b.k1 == 1.2 and np.array_equal(a.idf_np, b.idf_np) and np.array_equal(a.post_doc_np, b.post_doc_np)
    assert np.array_equal(a.post_val_np, b.post_val_np) and a.vocab == b.vocab
