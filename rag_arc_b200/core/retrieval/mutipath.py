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
        if self._canon_cache is not None and self._canon_cache[0] == sig:
            return self._canon_cache[1], self._canon_cache[2]
        keys: Dict[str, int] = {}
        tables, tabs = [], []
        for r in self.retrievers:
            docs = r.row_documents()
            tab = np.fromiter((keys.setdefault(d.content, len(keys)) for d in docs), np.int32, len(docs))
            tabs.append(tab)
            # device table with one extra trailing -1: indexing it with row -1 (padding) yields key -1
            tables.append(torch.from_numpy(np.concatenate([tab, np.array([-1], np.int32)])).to(device))
        # per retriever: key -> the LAST row holding that content (-1: not in this retriever's corpus)
        key_last_row = np.full((len(self.retrievers), max(len(keys), 1)), -1, np.int64)
        for l, tab in enumerate(tabs):
            key_last_row[l, tab] = np.arange(len(tab))       # later rows overwrite earlier ones
        # the same table on the device (resolution of fused keys to (retriever, row) happens there) and the
        # documents of every retriever as one object array (row -> Document by fancy indexing, plus a
        # trailing None that row -1 selects)
        key_last_row_dev = torch.from_numpy(key_last_row).to(device)
        doc_arrays = []
        for r in self.retrievers:
            docs = r.row_documents()
            arr = np.empty(len(docs) + 1, dtype=object)
            arr[:len(docs)] = docs
            arr[len(docs)] = None
            doc_arrays.append(arr)
        self._canon_cache = (sig, keys, tables, key_last_row, key_last_row_dev, doc_arrays)
        return keys, tables

    def invoke_batch(self, queries: List[str], **kwargs: Any) -> List[List[Document]]:
        top_k = kwargs.get("top_k", 10)
        if not all(hasattr(r, "batch_rows") for r in self.retrievers) or not isinstance(self.fusion_method, RRFusion):
            return [self.invoke(q, **kwargs) for q in queries]
        kl = self.top_k_per_retriever
        device = torch.device(getattr(self.fusion_method, "device", "cuda"))
        _, tables = self._canonical_tables(device)
        nq = len(queries)
        L = len(self.retrievers)
        per_list = []
        for l, r in enumerate(self.retrievers):
            keys = None
            try:
                if len(r.row_documents()) > 0:
                    rows = r.batch_rows(queries, kl).to(device)
                    # row -> content key through the device table; -1 rows (padding) stay -1: the table has a
                    # trailing -1 entry that row -1 indexes
                    keys = tables[l][rows]
                    if keys.shape[1] < kl:
                        keys = torch.nn.functional.pad(keys, (0, kl - keys.shape[1]), value=-1)
            except Exception as exc:  # noqa: BLE001
                print(f"retriever {type(r).__name__} failed: {exc}")
                keys = None
            per_list.append(keys if keys is not None else torch.full((nq, kl), -1, dtype=torch.int32, device=device))
        ids = torch.stack(per_list, 0)
        fused_ids, _, counts = self.fusion_method.fuse_batch(ids, top_k)
        # reference semantics: the Document returned for a content string is the LAST one seen
        # while walking the lists in retriever order (Fusion.py:61) - i.e. it comes from the last
        # retriever whose list holds the key; inside one retriever duplicate contents resolve to the
        # last row carrying them.  Resolved on the device; one small device->host transfer.
        key_last_row_dev, doc_arrays = self._canon_cache[4], self._canon_cache[5]
        present = (ids[:, :, None, :] == fused_ids[None, :, :, None]).any(dim=3)       # [L, nq, top_k]
        last_l = (L - 1) - torch.argmax(present.flip(0).to(torch.uint8), dim=0)        # [nq, top_k]
        rows = key_last_row_dev[last_l, fused_ids.clamp(min=0).long()]                 # [nq, top_k]
        rows = torch.where(fused_ids >= 0, rows, torch.full_like(rows, -1))
        packed = torch.cat([last_l.reshape(-1), rows.reshape(-1), counts.reshape(-1).long()]).cpu().numpy()
        last_l = packed[:nq * top_k].reshape(nq, top_k)
        rows = packed[nq * top_k:2 * nq * top_k].reshape(nq, top_k)
        counts = packed[2 * nq * top_k:]
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
