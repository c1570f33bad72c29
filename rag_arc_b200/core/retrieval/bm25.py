"""``BM25Retriever`` (drop-in for /root/reference core/retrieval/bm25.py) scored on the GPU.

Same surface as the reference class: ``from_texts`` / ``from_documents`` (:151-274),
``_get_relevant_documents`` with ``k = min(k, len(docs))`` (:276-317), ``get_scores`` (:319-333),
``get_top_k_with_scores`` (:335-366), add/delete with a full index rebuild and the same
RuntimeWarning threshold (:368-486), info/update helpers (:500-548), dill persistence (:550-609).
Construction keywords actually take effect here - the reference declares them as pydantic
``Field``s on a plain-ABC base, which silently drops them; the intended
behaviour (LangChain's BM25Retriever) is what is implemented.

The ``vectorizer`` is a ``B200BM25Okapi``: the rank_bm25 statistics are built on the host
(``Bm25Index``), scoring + top-k run in ``ragarc_bm25_topk`` / ``ragarc_bm25_scores``.
Ordering among equal scores is descending score, ascending document index (numpy's
``argsort(...)[::-1]`` order at :309 is implementation-defined).
"""
from __future__ import annotations

import asyncio
import functools
import logging
import os
import threading
import uuid
import warnings
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import dill
import numpy as np
import torch

from ... import ops
from ..utils.data_model import Document
from .base import BaseRetriever
from .bm25_index import Bm25Index

logger = logging.getLogger(__name__)


def default_preprocessing_func(text: str) -> List[str]:
    """Whitespace tokenisation (English only), the reference default (bm25.py:16-25)."""
    return text.split()


_load_device = threading.local()      # set by load_from_disk while it unpickles a natively saved vectorizer


class B200BM25Okapi:
    """Stand-in for ``rank_bm25.BM25Okapi``: same constructor and ``get_scores`` contract."""

    def __init__(self, corpus: Sequence[Sequence[str]], tokenizer=None, k1: float = 1.5, b: float = 0.75,
                 epsilon: float = 0.25, device="cuda"):
        if tokenizer is not None:
            corpus = [tokenizer(doc) for doc in corpus]
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.tokenizer = tokenizer
        self.index = Bm25Index.from_token_lists(corpus, k1=k1, b=b, epsilon=epsilon, device=device)

    corpus_size = property(lambda self: self.index.n_docs)
    avgdl = property(lambda self: self.index.avgdl)
    average_idf = property(lambda self: self.index.average_idf)
    doc_len = property(lambda self: self.index.doc_len_np.tolist())

    @property
    def idf(self) -> Dict[str, float]:
        return {tok: float(self.index.idf_np[ti]) for tok, ti in self.index.vocab.items()}

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        qt, ql = self.index.encode_queries([list(query)])
        return ops.bm25_scores(self.index, qt, ql)[0].cpu().numpy()

    def get_batch_topk(self, queries: Sequence[Sequence[str]], k: int):
        qt, ql = self.index.encode_queries(queries)
        return ops.bm25_topk(self.index, qt, ql, k)

    def get_batch_topk_texts(self, texts: Sequence[str], k: int):
        """Same for raw query strings under the default whitespace tokeniser: split + vocabulary lookup
        + packing happen in one library call (``ragarc_vocab_encode_split``), not per token in Python."""
        qt, ql = self.index.encode_texts(texts)
        return ops.bm25_topk(self.index, qt, ql, k)

    def __getstate__(self):
        ix = self.index
        return {"k1": self.k1, "b": self.b, "epsilon": self.epsilon, "tokenizer": self.tokenizer,
                "vocab": ix.vocab, "indptr": ix.indptr_np, "post_doc": ix.post_doc_np, "post_tf": ix.post_tf_np,
                "doc_len": ix.doc_len_np, "device": str(ix.device)}

    def __setstate__(self, st):
        self.k1, self.b, self.epsilon, self.tokenizer = st["k1"], st["b"], st["epsilon"], st["tokenizer"]
        # the device asked for by the loader (BM25Retriever.load_from_disk(device=...)) wins over the
        # one the state happened to be saved from
        device = getattr(_load_device, "value", None) or st["device"]
        self.index = Bm25Index(vocab=st["vocab"], indptr=st["indptr"], post_doc=st["post_doc"],
                               post_tf=st["post_tf"], doc_len=st["doc_len"], k1=self.k1, b=self.b,
                               epsilon=self.epsilon, device=device)


class BM25Retriever(BaseRetriever):
    def __init__(self, **kwargs: Any):
        warn_flag = kwargs.pop("warn_default_preprocess", True)
        self.vectorizer = kwargs.pop("vectorizer", None)
        self.docs: List[Document] = kwargs.pop("docs", None) or []
        k = kwargs.pop("k", 5)
        if not isinstance(k, int) or k <= 0:
            raise ValueError(f"k must be greater than 0, got {k}")
        self.k = k
        explicit_pre = "preprocess_func" in kwargs
        self.preprocess_func: Callable[[str], List[str]] = kwargs.pop("preprocess_func", default_preprocessing_func)
        if not callable(self.preprocess_func):
            raise ValueError("preprocess_func must be callable")
        self.bm25_params: Dict[str, Any] = kwargs.pop("bm25_params", None) or {}
        self.device = kwargs.pop("device", "cuda")
        self._version = 0            # bumped by every rebuild / add / delete (row -> Document caches key on it)
        super().__init__(**kwargs)
        if warn_flag and self.preprocess_func == default_preprocessing_func and not explicit_pre:
            warnings.warn("using the default whitespace tokenizer; provide preprocess_func for Chinese or other "
                          "languages", UserWarning, stacklevel=2)

    # ---- construction ----------------------------------------------------------------------------
    @classmethod
    def from_texts(cls, texts: Iterable[str], metadatas: Optional[Iterable[Dict[str, Any]]] = None,
                   ids: Optional[Iterable[str]] = None, bm25_params: Optional[Dict[str, Any]] = None,
                   preprocess_func: Callable[[str], List[str]] = default_preprocessing_func,
                   **kwargs: Any) -> "BM25Retriever":
        texts_list = list(texts)
        if not texts_list:
            raise ValueError("texts must not be empty")
        if metadatas is not None:
            metas = list(metadatas)
            if len(metas) != len(texts_list):
                raise ValueError(f"metadatas length ({len(metas)}) does not match texts length ({len(texts_list)})")
        else:
            metas = [{} for _ in texts_list]
        if ids is not None:
            ids_list = list(ids)
            if len(ids_list) != len(texts_list):
                raise ValueError(f"ids length ({len(ids_list)}) does not match texts length ({len(texts_list)})")
        else:
            ids_list = [str(uuid.uuid4()) for _ in texts_list]
        bm25_params = bm25_params or {}
        device = kwargs.get("device", "cuda")
        vectorizer = B200BM25Okapi([preprocess_func(t) for t in texts_list], device=device, **bm25_params)
        docs = [Document(content=t, metadata=m, id=i) for t, m, i in zip(texts_list, metas, ids_list)]
        return cls(vectorizer=vectorizer, docs=docs, preprocess_func=preprocess_func, bm25_params=bm25_params,
                   warn_default_preprocess=False, **kwargs)

    @classmethod
    def from_documents(cls, documents: Iterable[Document], bm25_params: Optional[Dict[str, Any]] = None,
                       preprocess_func: Callable[[str], List[str]] = default_preprocessing_func,
                       **kwargs: Any) -> "BM25Retriever":
        docs = list(documents)
        if not docs:
            raise ValueError("documents must not be empty")
        return cls.from_texts([d.content for d in docs], [d.metadata for d in docs], [d.id for d in docs],
                              bm25_params=bm25_params, preprocess_func=preprocess_func, **kwargs)

    def _rebuild(self) -> None:
        self._version += 1
        tokens = [self.preprocess_func(d.content) for d in self.docs]
        self.vectorizer = B200BM25Okapi(tokens, device=self.device, **self.bm25_params)

    # ---- search ----------------------------------------------------------------------------------
    def _get_relevant_documents(self, query: str, **kwargs: Any) -> List[Document]:
        if self.vectorizer is None:
            raise ValueError("BM25 vectorizer is not initialised")
        if not self.docs:
            logger.warning("document list is empty, returning no results")
            return []
        k = min(kwargs.get("k", self.k), len(self.docs))
        try:
            _, ids = self.vectorizer.get_batch_topk([self.preprocess_func(query)], k)
            return [self.docs[i] for i in ids[0].tolist() if i >= 0]
        except Exception as exc:
            logger.error("error during BM25 search: %s", exc)
            raise

    def get_scores(self, query: str) -> List[float]:
        if self.vectorizer is None:
            raise ValueError("BM25 vectorizer is not initialised")
        return self.vectorizer.get_scores(self.preprocess_func(query)).tolist()

    def get_top_k_with_scores(self, query: str, k: Optional[int] = None) -> List[Tuple[Document, float]]:
        if self.vectorizer is None:
            raise ValueError("BM25 vectorizer is not initialised")
        if not self.docs:
            return []
        k = min(k or self.k, len(self.docs))
        scores, ids = self.vectorizer.get_batch_topk([self.preprocess_func(query)], k)
        return [(self.docs[i], s) for i, s in zip(ids[0].tolist(), scores[0].tolist()) if i >= 0]

    # ---- batched (B200 addition) -----------------------------------------------------------------
    def search_batch(self, queries: List[str], k: Optional[int] = None):
        """-> (scores float64 [nq,k], doc indices int64 [nq,k]) device tensors, -1 padded."""
        if self.vectorizer is None:
            raise ValueError("BM25 vectorizer is not initialised")
        k = k or self.k
        if (self.preprocess_func is default_preprocessing_func and hasattr(self.vectorizer, "get_batch_topk_texts")
                and getattr(self.vectorizer, "tokenizer", None) is None and self.vectorizer.index.vocab is not None):
            return self.vectorizer.get_batch_topk_texts(queries, k)
        return self.vectorizer.get_batch_topk([self.preprocess_func(q) for q in queries], k)

    def batch_rows(self, queries: List[str], k: int):
        return self.search_batch(queries, min(k, max(len(self.docs), 1)))[1]

    def row_documents(self) -> List[Document]:
        return self.docs

    def corpus_stamp(self):
        """Changes whenever the row -> Document mapping may have changed (a delete followed by an add
        of equal size keeps ``len(docs)`` but moves rows): mutation counter + identity of the list."""
        return (self._version, id(self.docs), len(self.docs))

    def invoke_batch(self, queries: List[str], **kwargs: Any) -> List[List[Document]]:
        if not self.docs:
            return [[] for _ in queries]
        k = min(kwargs.get("k", self.k), len(self.docs))
        ids = self.search_batch(queries, k)[1].cpu().numpy()
        return [[self.docs[i] for i in row if i >= 0] for row in ids]

    # ---- maintenance -----------------------------------------------------------------------------
    def add_documents(self, documents: List[Document], **kwargs: Any) -> List[str]:
        if not documents:
            return []
        total = len(self.docs) + len(documents)
        if total > kwargs.get("rebuild_threshold", 1000):
            warnings.warn(f"rebuilding a BM25 index of {total} documents; this may be slow", RuntimeWarning,
                          stacklevel=2)
        self.docs.extend(documents)
        self._rebuild()
        return [d.id for d in documents if d.id is not None]

    async def aadd_documents(self, documents: List[Document], **kwargs: Any) -> List[str]:
        loop = asyncio.get_event_loop()
        with ThreadPoolExecutor() as pool:
            return await loop.run_in_executor(pool, functools.partial(self.add_documents, documents, **kwargs))

    def delete_documents(self, ids: Optional[List[str]] = None, **kwargs: Any) -> bool:
        self._version += 1
        if ids is None:
            self.docs.clear()
            self.vectorizer = None
            return True
        before = len(self.docs)
        self.docs = [d for d in self.docs if d.id not in ids]
        removed = before - len(self.docs)
        if removed > 0:
            if len(self.docs) > kwargs.get("rebuild_threshold", 1000):
                warnings.warn(f"rebuilding a BM25 index of {len(self.docs)} documents; this may be slow",
                              RuntimeWarning, stacklevel=2)
            if self.docs:
                self._rebuild()
            else:
                self.vectorizer = None
        return removed > 0

    async def adelete_documents(self, ids: Optional[List[str]] = None, **kwargs: Any) -> bool:
        loop = asyncio.get_event_loop()
        with ThreadPoolExecutor() as pool:
            return await loop.run_in_executor(pool, functools.partial(self.delete_documents, ids, **kwargs))

    def get_document_count(self) -> int:
        return len(self.docs)

    def get_bm25_info(self) -> Dict[str, Any]:
        info = {"document_count": len(self.docs), "k": self.k, "bm25_params": self.bm25_params,
                "preprocess_func": self.preprocess_func.__name__, "has_vectorizer": self.vectorizer is not None}
        if self.vectorizer is not None:
            info.update({"vocab_size": len(self.vectorizer.index.vocab),
                         "average_doc_length": getattr(self.vectorizer, "avgdl", "N/A")})
        return info

    def update_k(self, new_k: int) -> None:
        if new_k <= 0:
            raise ValueError(f"k must be greater than 0, got {new_k}")
        self.k = new_k

    def get_name(self) -> str:
        return "BM25Retriever"

    def __repr__(self) -> str:
        return f"{type(self).__name__}(docs={len(self.docs)}, k={self.k}, preprocess_func={self.preprocess_func.__name__})"

    # ---- persistence -----------------------------------------------------------------------------
    def save_to_disk(self, path: str) -> None:
        if not path.endswith(".pkl"):
            path = os.path.join(path, "bm25.pkl")
        try:
            state = {"vectorizer": self.vectorizer, "docs": self.docs, "k": self.k,
                     "preprocess_func": self.preprocess_func, "bm25_params": self.bm25_params}
            with open(path, "wb") as fh:
                dill.dump(state, fh)
        except Exception as exc:
            raise IOError(f"save failed: {exc}")

    @classmethod
    def load_from_disk(cls, path: str, device: Optional[str] = None) -> "BM25Retriever":
        """Loads a file written by ``save_to_disk`` - this class's or the REFERENCE's
        (core/retrieval/bm25.py:550-576).  A reference file carries a ``rank_bm25.BM25Okapi``; its
        parameters (k1, b, epsilon) are taken over and the CSR index is rebuilt from the documents
        with the stored tokeniser, which reproduces the reference's idf table bit for bit.
        ``device``: where the index goes; ``None`` = the device a natively saved state came from
        ("cuda" for a reference file)."""
        if not os.path.exists(path):
            raise IOError(f"file does not exist: {path}")
        try:
            from ...formats import ForeignBM25, load_reference_bm25_state
            _load_device.value = device
            try:
                st = load_reference_bm25_state(path)
            finally:
                _load_device.value = None
            vec = st["vectorizer"]
            pre = st.get("preprocess_func") or default_preprocessing_func
            params = dict(st.get("bm25_params") or {})
            if isinstance(vec, ForeignBM25):
                for name in ("k1", "b", "epsilon"):
                    if hasattr(vec, name):
                        params.setdefault(name, getattr(vec, name))
                vec = B200BM25Okapi([pre(d.content) for d in st["docs"]], device=device or "cuda", **params)
            return cls(vectorizer=vec, docs=st["docs"], k=st["k"], preprocess_func=pre,
                       bm25_params=st.get("bm25_params") or {}, warn_default_preprocess=False,
                       device=str(vec.index.device))
        except Exception as exc:
            raise IOError(f"load failed: {exc}")
