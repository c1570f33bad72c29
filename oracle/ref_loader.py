"""Import the reference's own hot-path Python as a live oracle (test infrastructure).

Only usable where /root/reference exists (this container) - never on the GPU box; everything
that must travel is exported by ``oracle/gen_golden.py`` into ``tests/golden/``.

The reference tree is read-only and partly broken as shipped, so:

* ``core/__init__.py:1-4`` imports a non-existent ``rag_arc`` package -> ``core`` is pre-seeded in
  ``sys.modules`` as a namespace stub pointing at the reference directory;
* ``core/retrieval/bm25.py:14`` imports ``utils.data_model`` -> a ``utils`` stub pointing at
  ``core/utils``;
* ``faiss`` / ``rank_bm25`` are satisfied by ``oracle/shims`` (numpy restatements);
* ``BM25Retriever`` declares pydantic ``Field`` attributes on a plain-ABC base
  (``bm25.py:82-90`` vs ``base.py:25-33``), so its attributes are assigned after construction.
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("RAGARC_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "core", "retrieval"))


def load():
    """Returns a namespace with the reference classes.  Idempotent."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REF}")
    sys.dont_write_bytecode = True
    for p in (_SHIMS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    # our own package must not shadow the reference's top-level names
    for name, sub in (("core", "core"), ("utils", "core/utils")):
        mod = sys.modules.get(name)
        if mod is None or getattr(mod, "__ragarc_ref_stub__", False) is False:
            stub = types.ModuleType(name)
            stub.__path__ = [os.path.join(REF, sub)]
            stub.__ragarc_ref_stub__ = True
            sys.modules[name] = stub
    ns = types.SimpleNamespace()
    from core.utils.Fusion import RRFusion, RetrievalResult, FusionMethod
    from core.utils.data_model import Document
    from core.retrieval.base import BaseRetriever
    from core.retrieval.mutipath import MultiPathRetriever
    from core.retrieval.dense import VectorStoreRetriever
    from core.file_management.embeddings.base import Embeddings
    from encapsulation.database.vector_db.VectorStoreBase import VectorStore
    from encapsulation.database.vector_db.VectorStore_Faiss import FaissVectorStore, _mmr_select
    from core.retrieval.bm25 import BM25Retriever
    from framework.register import Register
    from framework.config import AbstractConfig
    from framework.module import AbstractModule
    ns.__dict__.update(locals())
    return ns


def make_bm25_retriever(ns, texts, k=5, bm25_params=None, preprocess=str.split):
    """Build the reference ``BM25Retriever`` the way ``from_texts`` intends to
    (core/retrieval/bm25.py:151-238), working around the Field/ABC constructor bug."""
    from rank_bm25 import BM25Okapi
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = ns.BM25Retriever(warn_default_preprocess=False)
    r.vectorizer = BM25Okapi([preprocess(t) for t in texts], **(bm25_params or {}))
    r.docs = [ns.Document(content=t, metadata={}, id=str(i)) for i, t in enumerate(texts)]
    r.k = k
    r.preprocess_func = preprocess
    r.bm25_params = bm25_params or {}
    return r
