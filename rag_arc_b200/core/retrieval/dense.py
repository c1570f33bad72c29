"""``VectorStoreRetriever`` (drop-in for /root/reference core/retrieval/dense.py).

Search types ``similarity`` / ``similarity_score_threshold`` / ``mmr`` (:31-35), the same
validation errors (:61-84), keyword merge + default ``k=5`` + ``[:k]`` truncation (:122-174),
pass-through add/delete/get helpers (:220-330) and info helpers (:332-379).  ``invoke_batch`` is
the batched addition: it keeps the whole query batch on the device.
"""
from __future__ import annotations

import logging
from typing import Any, ClassVar, Collection, Dict, List, Optional

from ..utils.data_model import Document
from .base import BaseRetriever

logger = logging.getLogger(__name__)


class VectorStoreRetriever(BaseRetriever):
    allowed_search_types: ClassVar[Collection[str]] = ("similarity", "similarity_score_threshold", "mmr")

    def __init__(self, vectorstore, **kwargs: Any):
        self.vectorstore = vectorstore
        self.search_type = kwargs.get("search_type", "similarity")
        self.search_kwargs = kwargs.get("search_kwargs", {})
        self._validate_search_config()
        super().__init__(**kwargs)

    def _validate_search_config(self) -> None:
        if self.search_type not in self.allowed_search_types:
            raise ValueError(f"search_type '{self.search_type}' is not allowed; valid values: "
                             f"{self.allowed_search_types}")
        if self.search_type == "similarity_score_threshold":
            thr = self.search_kwargs.get("score_threshold")
            if thr is None or not isinstance(thr, (int, float)) or not (0 <= thr <= 1):
                raise ValueError("search_type 'similarity_score_threshold' needs a score_threshold in [0, 1] "
                                 "in search_kwargs")

    def _resolve(self, kwargs: Dict[str, Any]) -> Dict[str, Any]:
        params = {**self.search_kwargs, **kwargs}
        params["k"] = params.get("k", getattr(self, "k", 5))
        return params

    def _get_relevant_documents(self, query: str, **kwargs: Any) -> List[Document]:
        params = self._resolve(kwargs)
        k = params["k"]
        try:
            if self.search_type == "similarity":
                docs = self.vectorstore.similarity_search(query, **params)
            elif self.search_type == "similarity_score_threshold":
                docs = [d for d, _ in self.vectorstore.similarity_search_with_relevance_scores(query, **params)]
            elif self.search_type == "mmr":
                docs = self.vectorstore.max_marginal_relevance_search(query, **params)
            else:
                raise ValueError(f"unsupported search type: {self.search_type}")
            docs = docs[:k]
            logger.debug("retrieved %d documents (%s)", len(docs), self.search_type)
            return docs
        except Exception as exc:
            logger.error("error while retrieving documents: %s", exc)
            raise

    async def _aget_relevant_documents(self, query: str, **kwargs: Any) -> List[Document]:
        params = {**self.search_kwargs, **kwargs}
        try:
            if self.search_type == "similarity":
                return await self.vectorstore.asimilarity_search(query, **params)
            if self.search_type == "similarity_score_threshold":
                pairs = await self.vectorstore.asimilarity_search_with_relevance_scores(query, **params)
                return [d for d, _ in pairs]
            if self.search_type == "mmr":
                return await self.vectorstore.amax_marginal_relevance_search(query, **params)
            raise ValueError(f"unsupported search type: {self.search_type}")
        except Exception as exc:
            logger.error("error while retrieving documents asynchronously: %s", exc)
            raise

    # ---- batched (B200 addition) -----------------------------------------------------------------
    def batch_rows(self, queries: List[str], k: int):
        """Store rows of the top-k documents for every query: int64 ``[nq,k]`` on the device
        (-1 padded).  ``similarity`` search type only."""
        import numpy as np
        emb = self.vectorstore.embedding
        if hasattr(emb, "embed_documents_array"):      # our Embeddings base offers it; foreign plugins may not
            vecs = np.asarray(emb.embed_documents_array(list(queries)), dtype=np.float32)
        else:
            vecs = np.asarray(emb.embed_documents(list(queries)), dtype=np.float32)
        _, rows = self.vectorstore.search_batch(vecs, min(k, max(self.vectorstore.ntotal, 1)))
        return rows

    def row_documents(self) -> List[Document]:
        """Documents in store-row order (cached until the store changes: hybrid batches ask for it
        several times per call and the list is O(corpus))."""
        vs = self.vectorstore
        stamp = (getattr(vs, "_version", None), vs.ntotal)
        cached = getattr(self, "_row_docs", None)
        if cached is None or cached[0] != stamp or stamp[0] is None:
            cached = (stamp, [vs.docstore[vs.index_to_docstore_id[r]] for r in range(vs.ntotal)])
            self._row_docs = cached
        return cached[1]

    def corpus_stamp(self):
        """Changes whenever the store's row -> Document mapping may have changed."""
        vs = self.vectorstore
        return (getattr(vs, "_version", None), id(getattr(vs, "index_to_docstore_id", None)), vs.ntotal)

    def invoke_batch(self, queries: List[str], **kwargs: Any) -> List[List[Document]]:
        params = self._resolve(kwargs)
        if self.search_type != "similarity" or not hasattr(self.vectorstore, "search_batch"):
            return [self.invoke(q, **kwargs) for q in queries]
        if self.vectorstore.ntotal == 0:
            return [[] for _ in queries]
        rows = self.batch_rows(queries, params["k"]).cpu().numpy()
        return [self.vectorstore.rows_to_documents(r) for r in rows]

    # ---- pass-through helpers --------------------------------------------------------------------
    def add_documents(self, documents: List[Document], **kwargs: Any) -> List[str]:
        ids = self.vectorstore.add_documents(documents, **kwargs)
        logger.info("added %d documents to the vector store", len(documents))
        return ids

    async def aadd_documents(self, documents: List[Document], **kwargs: Any) -> List[str]:
        return await self.vectorstore.aadd_documents(documents, **kwargs)

    def delete_documents(self, ids: Optional[List[str]] = None, **kwargs: Any) -> Optional[bool]:
        return self.vectorstore.delete(ids, **kwargs)

    async def adelete_documents(self, ids: Optional[List[str]] = None, **kwargs: Any) -> Optional[bool]:
        return await self.vectorstore.adelete(ids, **kwargs)

    def get_by_ids(self, ids: List[str]) -> List[Document]:
        return self.vectorstore.get_by_ids(ids)

    async def aget_by_ids(self, ids: List[str]) -> List[Document]:
        return await self.vectorstore.aget_by_ids(ids)

    def get_vectorstore_info(self) -> Dict[str, Any]:
        info = {"vectorstore_class": type(self.vectorstore).__name__, "search_type": self.search_type,
                "search_kwargs": self.search_kwargs, "allowed_search_types": list(self.allowed_search_types)}
        emb = getattr(self.vectorstore, "embeddings", None) or getattr(self.vectorstore, "embedding", None)
        if emb is not None:
            info["embedding_class"] = type(emb).__name__
        return info

    def get_name(self) -> str:
        return f"{type(self.vectorstore).__name__}Retriever"

    def update_search_params(self, **kwargs: Any) -> None:
        self.search_kwargs.update(kwargs)
        if "search_type" in kwargs:
            self.search_type = kwargs["search_type"]
            self._validate_search_config()

    def __repr__(self) -> str:
        return (f"{type(self).__name__}(vectorstore={type(self.vectorstore).__name__}, "
                f"search_type='{self.search_type}', search_kwargs={self.search_kwargs})")
