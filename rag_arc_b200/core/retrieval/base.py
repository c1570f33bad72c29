"""Retriever plugin base (drop-in for /root/reference core/retrieval/base.py).

``invoke`` -> ``_get_relevant_documents`` (base.py:35-50,69-80); ``ainvoke`` runs the synchronous
method in a fresh thread pool unless a subclass overrides ``_aget_relevant_documents``
(base.py:52-67,82-96) - so the C extension may be entered from arbitrary threads.
"""
from __future__ import annotations

import asyncio
import functools
from abc import ABC, abstractmethod
from concurrent.futures import ThreadPoolExecutor
from typing import Any, List

from ..utils.data_model import Document


class BaseRetriever(ABC):
    def __init__(self, **kwargs: Any):
        self.search_kwargs = kwargs.get("search_kwargs", {})
        self.tags = kwargs.get("tags")
        self.metadata = kwargs.get("metadata")

    def invoke(self, input: str, **kwargs: Any) -> List[Document]:
        return self._get_relevant_documents(input, **kwargs)

    async def ainvoke(self, input: str, **kwargs: Any) -> List[Document]:
        return await self._aget_relevant_documents(input, **kwargs)

    @abstractmethod
    def _get_relevant_documents(self, query: str, **kwargs: Any) -> List[Document]:
        ...

    async def _aget_relevant_documents(self, query: str, **kwargs: Any) -> List[Document]:
        loop = asyncio.get_event_loop()
        with ThreadPoolExecutor() as pool:
            call = functools.partial(self._get_relevant_documents, query, **kwargs)
            return await loop.run_in_executor(pool, call)

    def invoke_batch(self, queries: List[str], **kwargs: Any) -> List[List[Document]]:
        """Batched form (B200 addition; the reference takes one query per call).  The default walks
        the queries one by one; the dense, BM25 and multi-path retrievers override it with one GPU
        batch."""
        return [self.invoke(q, **kwargs) for q in queries]

    def get_name(self) -> str:
        return type(self).__name__
