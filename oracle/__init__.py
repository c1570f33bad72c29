"""CPU oracle for the RAG-ARC retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``rag_arc_b200`` (the product) may import this package.  The only legal
importers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``, and there only as the checker or as the
reported CPU baseline - never as the thing that is shipped or measured as "ours".

What is restated here, and what pins it (citations relative to /root/reference):

* dense  (``oracle/dense.py``)  - ``faiss.normalize_L2`` + ``faiss.IndexFlatIP.search`` as
  called from ``encapsulation/database/vector_db/VectorStore_Faiss.py:150-154,258-263``.
  FAISS itself is an un-vendored, un-pinned third-party dependency (requirements.txt:1 lists
  only ``dill``) and is not installable here: **parity unpinned** for the fp32 summation order
  and tie order; the restatement follows FAISS's published flat-IP semantics (exact fp32 inner
  product, descending, -1 padding) and is pinned only by this repo's hand-computed KATs.
* BM25   (``oracle/bm25.py``)   - ``rank_bm25.BM25Okapi`` (un-vendored, un-pinned; published
  0.2.x algorithm) as called from ``core/retrieval/bm25.py:213-218,302-311``: **parity
  unpinned** w.r.t. the third-party package, pinned by hand-computed KATs.
* RRF / hybrid (``oracle/rrf.py``) - integer restatement of ``core/utils/Fusion.py:45-76``;
  **pinned**: checked against the reference's own ``RRFusion`` / ``MultiPathRetriever`` executed
  live (``oracle/ref_loader.py``), outputs committed under ``tests/golden/``.
* pool + normalise (``oracle/pool.py``) - sentence-transformers ``Pooling``/``Normalize``
  semantics behind ``core/file_management/embeddings/huggingface.py:122-134``: third-party,
  **parity unpinned**, pinned by KATs.
"""
