"""Vector-store plugin interface (drop-in for /root/reference
encapsulation/database/vector_db/VectorStoreBase.py).

Same template methods and error behaviour: ``search`` dispatch on search type (:184-212),
relevance-score maps ``1-d/sqrt(2)``, ``1-d`` and the piecewise inner-product map (:258-273),
``similarity_search_with_relevance_scores`` with the [0,1] warning and ``score_threshold``
filter (:347-392), ``from_documents`` (:526-552) and thread-pool async twins.  Two defects of the
reference are not reproduced: ``add_texts`` no longer dies on the TYPE_CHECKING-only ``Document``
import (:23-27,91) and ``as_retriever`` imports the retriever that exists (:625).
"""
from __future__ import annotations

import asyncio
import functools
import logging
import math
import warnings
from abc import ABC, abstractmethod
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from itertools import cycle
from typing import Any, Callable, Iterable, List, Optional, Sequence, Tuple

from ....core.utils.data_model import Document

logger = logging.getLogger(__name__)


@dataclass
class SearchResult:
    document: Document
    score: float
    distance: float


async def _in_thread(fn, *args, **kwargs):
    loop = asyncio.get_event_loop()
    return await loop.run_in_executor(ThreadPoolExecutor(), functools.partial(fn, *args, **kwargs))


class VectorStore(ABC):
    def __init__(self, **kwargs: Any):
        pass

    # ---- ingestion ---------------------------------------------------------------------------
    def add_texts(self, texts: Iterable[str], metadatas: Optional[List[dict]] = None, *,
                  ids: Optional[List[str]] = None, **kwargs: Any) -> List[str]:
        texts_ = texts if isinstance(texts, (list, tuple)) else list(texts)
        if metadatas and len(metadatas) != len(texts_):
            raise ValueError(f"number of metadatas ({len(metadatas)}) must match number of texts ({len(texts_)})")
        metas = iter(metadatas) if metadatas else cycle([{}])
        ids_ = iter(ids) if ids else cycle([None])
        docs = [Document(id=i, content=t, metadata=m) for t, m, i in zip(texts_, metas, ids_)]
        if ids is not None:
            kwargs["ids"] = ids
        return self.add_documents(docs, **kwargs)

    def add_documents(self, documents: List[Document], **kwargs: Any) -> List[str]:
        if "ids" not in kwargs:
            ids = [doc.id for doc in documents]
            if any(ids):
                kwargs["ids"] = ids
        return self.add_texts([d.content for d in documents], [d.metadata for d in documents], **kwargs)

    async def aadd_texts(self, texts, metadatas=None, *, ids=None, **kwargs):
        return await _in_thread(self.add_texts, texts, metadatas, ids=ids, **kwargs)

    async def aadd_documents(self, documents, **kwargs):
        return await _in_thread(self.add_documents, documents, **kwargs)

    def delete(self, ids: Optional[List[str]] = None, **kwargs: Any) -> Optional[bool]:
        raise NotImplementedError("delete must be implemented by the subclass")

    async def adelete(self, ids=None, **kwargs):
        return await _in_thread(self.delete, ids, **kwargs)

    def get_by_ids(self, ids: Sequence[str], /) -> List[Document]:
        raise NotImplementedError(f"{type(self).__name__} does not support get_by_ids yet")

    async def aget_by_ids(self, ids, /):
        return await _in_thread(self.get_by_ids, ids)

    # ---- search ------------------------------------------------------------------------------
    def search(self, query: str, search_type: str, **kwargs: Any) -> List[Document]:
        if search_type == "similarity":
            return self.similarity_search(query, **kwargs)
        if search_type == "similarity_score_threshold":
            return [doc for doc, _ in self.similarity_search_with_relevance_scores(query, **kwargs)]
        if search_type == "mmr":
            return self.max_marginal_relevance_search(query, **kwargs)
        raise ValueError(f"search_type {search_type} is not allowed; expected 'similarity', "
                         "'similarity_score_threshold' or 'mmr'")

    async def asearch(self, query: str, search_type: str, **kwargs: Any) -> List[Document]:
        if search_type == "similarity":
            return await self.asimilarity_search(query, **kwargs)
        if search_type == "similarity_score_threshold":
            return [doc for doc, _ in await self.asimilarity_search_with_relevance_scores(query, **kwargs)]
        if search_type == "mmr":
            return await self.amax_marginal_relevance_search(query, **kwargs)
        raise ValueError(f"search_type {search_type} is not allowed; expected 'similarity', "
                         "'similarity_score_threshold' or 'mmr'")

    @abstractmethod
    def similarity_search(self, query: str, k: int = 4, **kwargs: Any) -> List[Document]:
        ...

    async def asimilarity_search(self, query, k=4, **kwargs):
        return await _in_thread(self.similarity_search, query, k, **kwargs)

    @staticmethod
    def _euclidean_relevance_score_fn(distance: float) -> float:
        return 1.0 - distance / math.sqrt(2)

    @staticmethod
    def _cosine_relevance_score_fn(distance: float) -> float:
        return 1.0 - distance

    @staticmethod
    def _max_inner_product_relevance_score_fn(distance: float) -> float:
        if distance > 0:
            return 1.0 - distance
        return -1.0 * distance

    def _select_relevance_score_fn(self) -> Callable[[float], float]:
        raise NotImplementedError

    def similarity_search_with_score(self, *args: Any, **kwargs: Any) -> List[Tuple[Document, float]]:
        raise NotImplementedError

    async def asimilarity_search_with_score(self, *args, **kwargs):
        return await _in_thread(self.similarity_search_with_score, *args, **kwargs)

    def _similarity_search_with_relevance_scores(self, query: str, k: int = 4, **kwargs: Any):
        fn = self._select_relevance_score_fn()
        return [(doc, fn(score)) for doc, score in self.similarity_search_with_score(query, k, **kwargs)]

    async def _asimilarity_search_with_relevance_scores(self, query: str, k: int = 4, **kwargs: Any):
        fn = self._select_relevance_score_fn()
        return [(doc, fn(score)) for doc, score in await self.asimilarity_search_with_score(query, k, **kwargs)]

    @staticmethod
    def _apply_threshold(pairs, score_threshold):
        if any(s < 0.0 or s > 1.0 for _, s in pairs):
            warnings.warn(f"relevance scores must be between 0 and 1, got {pairs}", stacklevel=3)
        if score_threshold is not None:
            pairs = [(d, s) for d, s in pairs if s >= score_threshold]
            if not pairs:
                logger.warning("no relevant documents retrieved with relevance score threshold %s", score_threshold)
        return pairs

    def similarity_search_with_relevance_scores(self, query: str, k: int = 4, **kwargs: Any):
        score_threshold = kwargs.pop("score_threshold", None)
        return self._apply_threshold(self._similarity_search_with_relevance_scores(query, k=k, **kwargs),
                                     score_threshold)

    async def asimilarity_search_with_relevance_scores(self, query: str, k: int = 4, **kwargs: Any):
        score_threshold = kwargs.pop("score_threshold", None)
        return self._apply_threshold(await self._asimilarity_search_with_relevance_scores(query, k=k, **kwargs),
                                     score_threshold)

    def similarity_search_by_vector(self, embedding: List[float], k: int = 4, **kwargs: Any) -> List[Document]:
        raise NotImplementedError

    async def asimilarity_search_by_vector(self, embedding, k=4, **kwargs):
        return await _in_thread(self.similarity_search_by_vector, embedding, k, **kwargs)

    def max_marginal_relevance_search(self, query: str, k: int = 4, fetch_k: int = 20,
                                      lambda_mult: float = 0.5, **kwargs: Any) -> List[Document]:
        raise NotImplementedError

    async def amax_marginal_relevance_search(self, query, k=4, fetch_k=20, lambda_mult=0.5, **kwargs):
        return await _in_thread(self.max_marginal_relevance_search, query, k, fetch_k, lambda_mult, **kwargs)

    def max_marginal_relevance_search_by_vector(self, embedding: List[float], k: int = 4, fetch_k: int = 20,
                                                lambda_mult: float = 0.5, **kwargs: Any) -> List[Document]:
        raise NotImplementedError

    async def amax_marginal_relevance_search_by_vector(self, embedding, k=4, fetch_k=20, lambda_mult=0.5, **kwargs):
        return await _in_thread(self.max_marginal_relevance_search_by_vector, embedding, k, fetch_k,
                                lambda_mult, **kwargs)

    # ---- construction ------------------------------------------------------------------------
    @classmethod
    def from_documents(cls, documents: List[Document], embedding, **kwargs: Any) -> "VectorStore":
        if "ids" not in kwargs:
            ids = [doc.id for doc in documents]
            if any(ids):
                kwargs["ids"] = ids
        return cls.from_texts([d.content for d in documents], embedding,
                              metadatas=[d.metadata for d in documents], **kwargs)

    @classmethod
    async def afrom_documents(cls, documents: List[Document], embedding, **kwargs: Any) -> "VectorStore":
        if "ids" not in kwargs:
            ids = [doc.id for doc in documents]
            if any(ids):
                kwargs["ids"] = ids
        return await cls.afrom_texts([d.content for d in documents], embedding,
                                     metadatas=[d.metadata for d in documents], **kwargs)

    @classmethod
    @abstractmethod
    def from_texts(cls, texts: List[str], embedding, metadatas: Optional[List[dict]] = None, *,
                   ids: Optional[List[str]] = None, **kwargs: Any) -> "VectorStore":
        ...

    @classmethod
    async def afrom_texts(cls, texts, embedding, metadatas=None, *, ids=None, **kwargs):
        if ids is not None:
            kwargs["ids"] = ids
        return await _in_thread(cls.from_texts, texts, embedding, metadatas, **kwargs)

    def _get_retriever_tags(self) -> List[str]:
        tags = [type(self).__name__]
        emb = getattr(self, "embeddings", None) or getattr(self, "embedding", None)
        if emb is not None:
            tags.append(type(emb).__name__)
        return tags

    def as_retriever(self, **kwargs: Any):
        from ....core.retrieval.dense import VectorStoreRetriever
        tags = (kwargs.pop("tags", None) or []) + self._get_retriever_tags()
        return VectorStoreRetriever(vectorstore=self, tags=tags, **kwargs)
