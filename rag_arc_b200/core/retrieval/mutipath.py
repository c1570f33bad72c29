"""``MultiPathRetriever`` (drop-in for /root/reference core/retrieval/mutipath.py).

Single-query path identical in behaviour to the reference (:37-93): retrievers are invoked in
order with ``k = top_k_per_retriever`` forced (other kwargs, ``top_k`` included, are forwarded),
a retriever that raises contributes an empty list and a printed message (:78-80), all-empty gives
``[]`` (:83-84), results are wrapped as ``RetrievalResult(rank=i+1)`` and fused with
``fusion_method.fuse(all_results, top_k)`` (``top_k`` default 10, :56), documents only are returned.

``invoke_batch`` is the B200 addition: every retriever that can answer a whole query batch on
the device (``batch_rows``) does so, rows are translated to content-canonical integer keys with a
device lookup table and the whole batch is fused by one ``ragarc_rrf_fuse`` launch.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np
import torch

from ..utils.data_model import Document
from ..utils.Fusion import FusionMethod, RetrievalResult, RRFusion
from .base import BaseRetriever


class MultiPathRetriever(BaseRetriever):
    def __init__(self, retrievers: List[BaseRetriever], fusion_method: Optional[FusionMethod] = None,
                 top_k_per_retriever: int = 50):
        self.retrievers = retrievers
        self.fusion_method = fusion_method or RRFusion()
        self.top_k_per_retriever = top_k_per_retriever
        self._canon_cache = None

    def _get_relevant_documents(self, query: str, **kwargs: Any) -> List[Document]:
        top_k = kwargs.get("top_k", 10)
        all_results: List[List[RetrievalResult]] = []
        for retriever in self.retrievers:
            try:
                docs = retriever.invoke(query, **{**kwargs, "k": self.top_k_per_retriever})
                all_results.append([RetrievalResult(document=d, score=getattr(d, "score", 1.0), rank=i + 1)
                                    for i, d in enumerate(docs)])
            except Exception as exc:  # noqa: BLE001 - reference behaviour: swallow, report, continue
                print(f"retriever {type(retriever).__name__} failed: {exc}")
                all_results.append([])
        if not all_results or all(len(r) == 0 for r in all_results):
            return []
        return [r.document for r in self.fusion_method.fuse(all_results, top_k)]

    # ---- batched hybrid (B200 addition) ------------------------------------------------------------
    def _canonical_tables(self, device):
        """content -> integer key (first appearance over the retrievers, in order) and, per
        retriever, a device table row -> key.  Rebuilt when any retriever reports a different
        ``corpus_stamp()`` (mutation counter, not just the corpus size: a delete followed by an add of
        the same number of documents moves rows without changing the size)."""
        sig = tuple((id(r), r.corpus_stamp()) if hasattr(r, "corpus_stamp")
                    else (id(r), id(r.row_documents()), len(r.row_documents())) for r in self.retrievers)
        cache = self._canon_cache                  # one read: another thread may replace it meanwhile
        if cache is not None and cache[0] == sig:
            return cache[1], cache[2], cache[3]
        keys: Dict[str, int] = {}
        tables, doc_arrays = [], []
        for r in self.retrievers:
            docs = r.row_documents()
            tab = np.fromiter((keys.setdefault(d.content, len(keys)) for d in docs), np.int32, len(docs))
            tables.append(torch.from_numpy(tab).to(device))
            # the documents of the retriever as one object array: row -> Document by fancy indexing, plus a
            # trailing None that row -1 selects
            arr = np.empty(len(docs) + 1, dtype=object)
            arr[:len(docs)] = docs
            arr[len(docs)] = None
            doc_arrays.append(arr)
        self._canon_cache = (sig, keys, tables, doc_arrays)
        return keys, tables, doc_arrays

    def invoke_batch(self, queries: List[str], **kwargs: Any) -> List[List[Document]]:
        top_k = kwargs.get("top_k", 10)
        if not all(hasattr(r, "batch_rows") for r in self.retrievers) or not isinstance(self.fusion_method, RRFusion):
            return [self.invoke(q, **kwargs) for q in queries]
        kl = self.top_k_per_retriever
        device = torch.device(getattr(self.fusion_method, "device", "cuda"))
        _, tables, doc_arrays = self._canonical_tables(device)
        nq = len(queries)
        per_list = []
        for l, r in enumerate(self.retrievers):
            rows = None
            try:
                if len(r.row_documents()) > 0:
                    rows = r.batch_rows(queries, kl).to(device)
            except Exception as exc:  # noqa: BLE001
                print(f"retriever {type(r).__name__} failed: {exc}")
                rows = None
            per_list.append(rows)
        if all(rows is None for rows in per_list):
            return [[] for _ in queries]
        # one launch: row -> content key through the device tables, RRF, and for every fused key the
        # (retriever, row) of the Document the reference returns for it - the LAST one seen while walking the
        # lists in retriever order (document_map[content] is overwritten, Fusion.py:61).  One small
        # device->host transfer brings (list, row, count) back.
        _, _, packed = self.fusion_method.fuse_rows_batch(per_list, tables, kl, top_k)
        from ... import ops
        last_l, rows, counts = ops.unpack_fused_rows(packed.cpu().numpy(), nq, top_k)
        # row -> Document by fancy indexing into the per-retriever object arrays (row -1 -> None)
        out = np.empty((nq, top_k), dtype=object)
        for l, arr in enumerate(doc_arrays):
            sel = last_l == l
            if sel.any():
                out[sel] = arr[rows[sel]]
        full = out.tolist()
        if int(counts.min(initial=top_k)) >= top_k:
            return full
        return [row[:c] for row, c in zip(full, counts.tolist())]

    # ---- management --------------------------------------------------------------------------------
    def add_retriever(self, retriever: BaseRetriever):
        self.retrievers.append(retriever)
        self._canon_cache = None

    def remove_retriever(self, name: str):
        for i, r in enumerate(self.retrievers):
            if type(r).__name__ == name:
                self.retrievers.pop(i)
                break
        self._canon_cache = None

    def set_fusion_method(self, fusion_method: FusionMethod):
        self.fusion_method = fusion_method
