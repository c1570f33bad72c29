"""rag_arc_b200 - B200 (sm_100a) implementation of RAG-ARC's retrieval hot path.

Exact dense top-k (tcgen05 tensor-core scoring fused with per-query selection), BM25 over CSR
postings, reciprocal-rank fusion and pool+normalise, behind RAG-ARC's own plugin surface
(``VectorStore`` / ``BaseRetriever`` / ``FusionMethod`` / ``Embeddings`` / ``AbstractConfig``).
All arithmetic runs in ``libragarc_b200.so`` (hand-written CUDA behind the C ABI in
``include/ragarc_b200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"
