"""Embeddings plugin interface: the two abstract methods and the two thread-pool async wrappers of
/root/reference core/file_management/embeddings/base.py:7-61, plus one batched accessor
(``embed_documents_array``) that the B200 retrievers use so that a batch of query vectors reaches
the device as one array instead of a list of Python float lists."""
from __future__ import annotations

import asyncio
from abc import ABC, abstractmethod
from concurrent.futures import ThreadPoolExecutor
from typing import List

import numpy as np

_POOL = ThreadPoolExecutor(thread_name_prefix="embeddings")


def _in_pool(fn, arg):
    return asyncio.get_event_loop().run_in_executor(_POOL, fn, arg)


class Embeddings(ABC):
    def __init__(self, **kwargs):
        # the reference's constructor accepts and ignores arbitrary keyword arguments
        pass

    @abstractmethod
    def embed_documents(self, texts: List[str]) -> List[List[float]]:
        """One embedding (list of floats) per input text."""

    @abstractmethod
    def embed_query(self, text: str) -> List[float]:
        """The embedding of a single query string."""

    def embed_documents_array(self, texts: List[str]) -> np.ndarray:
        """``embed_documents`` as one float32 ``[n, d]`` array.  Subclasses that already hold their
        vectors as arrays or tensors override this to skip the list-of-floats round trip."""
        return np.asarray(self.embed_documents(list(texts)), dtype=np.float32)

    async def aembed_documents(self, texts: List[str]) -> List[List[float]]:
        return await _in_pool(self.embed_documents, texts)

    async def aembed_query(self, text: str) -> List[float]:
        return await _in_pool(self.embed_query, text)
