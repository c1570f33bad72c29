"""Result fusion plugins (drop-in for /root/reference core/utils/Fusion.py).

``RRFusion.fuse`` keeps the reference contract (Fusion.py:45-76): ranks are re-assigned from list
position, documents are de-duplicated by *content*, ``document_map`` keeps the last document seen
for a content string, ordering is score-descending with first-appearance tie order, and the
returned ``RetrievalResult`` carry the fused fp64 score and rank ``1..``.  The arithmetic (rank ->
``1.0/(k+rank)``, accumulation, ordering) runs in the ``ragarc_rrf_fuse`` kernel on integer keys;
``fuse_batch`` is the batched entry the hybrid retriever uses to keep a whole query batch on device.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np
import torch

from .data_model import Document


@dataclass
class RetrievalResult:
    document: Document
    score: float
    rank: int = 0


class FusionMethod(ABC):
    @abstractmethod
    def fuse(self, results: List[List[RetrievalResult]], top_k: int) -> List[RetrievalResult]:
        """Merge the ranked lists of several retrievers into one ranked list of ``top_k``."""


class RRFusion(FusionMethod):
    def __init__(self, k: float = 60.0, device="cuda"):
        self.k = k
        self.device = device

    def fuse_batch(self, ids: torch.Tensor, top_k: int):
        """ids: int32 [L, nq, kl] canonical document keys (negative = padding), already on the GPU.
        Returns (ids int32 [nq, top_k], scores float64 [nq, top_k], count int32 [nq])."""
        from ... import ops
        return ops.rrf_fuse(ids, top_k, float(self.k))

    def fuse_rows_batch(self, rows, row_to_key, kl: int, top_k: int):
        """The batched hybrid merge on retriever rows (ops.rrf_fuse_rows): fusion plus the (list, row) of the
        Document the reference's ``document_map`` would hold for every fused key, in one launch."""
        from ... import ops
        return ops.rrf_fuse_rows(rows, row_to_key, kl, top_k, float(self.k))

    def fuse(self, results: List[List[RetrievalResult]], top_k: int) -> List[RetrievalResult]:
        for ranked in results:                       # Fusion.py:47-49
            for pos, res in enumerate(ranked):
                res.rank = pos + 1
        keys: Dict[str, int] = {}
        last_doc: Dict[int, Document] = {}
        kl = max((len(r) for r in results), default=0)
        if kl == 0 or top_k <= 0:
            return []
        arr = np.full((len(results), 1, kl), -1, np.int32)
        for l, ranked in enumerate(results):
            for pos, res in enumerate(ranked):
                key = keys.setdefault(res.document.content, len(keys))
                last_doc[key] = res.document          # Fusion.py:61 keeps the last one seen
                arr[l, 0, pos] = key
        from ... import ops
        dev_ids = torch.from_numpy(arr).to(self.device)
        out_ids, out_scores, out_count = ops.rrf_fuse(dev_ids, int(top_k), float(self.k))
        n = int(out_count[0])
        ids_h = out_ids[0, :n].tolist()
        sc_h = out_scores[0, :n].tolist()
        return [RetrievalResult(document=last_doc[key], score=score, rank=pos + 1)
                for pos, (key, score) in enumerate(zip(ids_h, sc_h))]
