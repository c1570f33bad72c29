"""Oracle (test infrastructure): flat inner-product search as the reference calls it.

Restates what ``FaissVectorStore`` obtains from FAISS at
``encapsulation/database/vector_db/VectorStore_Faiss.py``:

* ``:150-154``  ``faiss.normalize_L2(vectors)`` when metric is cosine / normalize_L2 is set
* ``:169-178,202``  ``np.float32`` rows appended with ``index.add``
* ``:258-263``  ``index.search(q[1,d], min(k, ntotal))`` -> ``(D float32[nq,k], I int64[nq,k])``
* ``:267``  ``-1`` ids are skipped

FAISS is third-party and absent from /root/reference (**parity unpinned**, see package
docstring); this file restates its published flat-index semantics with numpy.
Tie rule of the oracle: equal scores are ordered by ascending row id (FAISS leaves it
unspecified; comparators in ``oracle/compare.py`` are tie-aware).
"""
from __future__ import annotations

import numpy as np

__all__ = ["normalize_L2", "IndexFlatIP", "IndexFlatL2", "flat_ip_search", "flat_ip_search_f64",
           "topk_desc_stable"]


def normalize_L2(x: np.ndarray) -> None:
    """In place, per row: ``nr = sum(x*x)`` in fp32; if ``nr > 0``: ``x *= 1/sqrt(nr)``.

    FAISS semantics used at VectorStore_Faiss.py:153 (zero rows are left untouched, no eps).
    """
    assert x.dtype == np.float32 and x.ndim == 2
    nr = np.einsum("ij,ij->i", x, x, dtype=np.float32)
    nz = nr > 0
    inv = np.ones_like(nr)
    inv[nz] = np.float32(1.0) / np.sqrt(nr[nz], dtype=np.float32)
    x *= inv[:, None]


def topk_desc_stable(scores: np.ndarray, k: int):
    """Exact top-k of each row, descending score, ties by ascending column id.

    Returns (D[nq,k] same dtype as scores, I[nq,k] int64); pads with (-inf, -1) when k > n.
    """
    nq, n = scores.shape
    kk = min(k, n)
    D = np.full((nq, k), -np.inf, dtype=scores.dtype)
    I = np.full((nq, k), -1, dtype=np.int64)
    if kk == 0:
        return D, I
    for i in range(nq):
        row = scores[i]
        if kk < n:
            # all elements >= the kk-th largest value (keeps every tie of the boundary value)
            kth = np.partition(row, n - kk)[n - kk]
            cand = np.nonzero(row >= kth)[0]
        else:
            cand = np.arange(n)
        # lexsort: last key is primary -> sort by (-score, id)
        order = np.lexsort((cand, -row[cand].astype(np.float64)))[:kk]
        sel = cand[order]
        D[i, :kk] = row[sel]
        I[i, :kk] = sel
    return D, I


def flat_ip_search(X32: np.ndarray, Q32: np.ndarray, k: int, block: int = 65536):
    """``IndexFlatIP.search``: fp32 inner products of every query with every row + exact top-k.

    Mirrors the call at VectorStore_Faiss.py:263.  Blocked over the corpus so that 1M-row
    corpora do not materialise the full score matrix at once.
    """
    assert X32.dtype == np.float32 and Q32.dtype == np.float32
    nq = Q32.shape[0]
    n = X32.shape[0]
    bestD = np.full((nq, 0), 0, dtype=np.float32)
    bestI = np.full((nq, 0), 0, dtype=np.int64)
    for s in range(0, n, block):
        S = Q32 @ X32[s:s + block].T
        D, I = topk_desc_stable(S, min(k, S.shape[1]))
        I = I + s
        bestD = np.concatenate([bestD, D], axis=1)
        bestI = np.concatenate([bestI, I], axis=1)
        if bestD.shape[1] > k:
            bestD, bestI = _merge(bestD, bestI, k)
    if bestD.shape[1] < k:
        pad = k - bestD.shape[1]
        bestD = np.concatenate([bestD, np.full((nq, pad), -np.inf, np.float32)], axis=1)
        bestI = np.concatenate([bestI, np.full((nq, pad), -1, np.int64)], axis=1)
    else:
        bestD, bestI = _merge(bestD, bestI, k)
    return bestD, bestI


def _merge(D, I, k):
    nq = D.shape[0]
    oD = np.empty((nq, k), D.dtype)
    oI = np.empty((nq, k), np.int64)
    for i in range(nq):
        order = np.lexsort((I[i], -D[i].astype(np.float64)))[:k]
        oD[i] = D[i][order]
        oI[i] = I[i][order]
    return oD, oI


def flat_ip_search_f64(X: np.ndarray, Q: np.ndarray, k: int, block: int = 65536):
    """Adjudicator: same search with fp64 accumulation (inputs are the storage-dtype values
    upcast exactly).  Used to decide whether an id swap sits inside a numerical tie group."""
    nq = Q.shape[0]
    n = X.shape[0]
    Q64 = Q.astype(np.float64)
    bestD = np.full((nq, 0), 0, dtype=np.float64)
    bestI = np.full((nq, 0), 0, dtype=np.int64)
    for s in range(0, n, block):
        S = Q64 @ X[s:s + block].astype(np.float64).T
        D, I = topk_desc_stable(S, min(k, S.shape[1]))
        bestD = np.concatenate([bestD, D], axis=1)
        bestI = np.concatenate([bestI, I + s], axis=1)
        if bestD.shape[1] > k:
            bestD, bestI = _merge(bestD, bestI, k)
    if bestD.shape[1] >= k:
        bestD, bestI = _merge(bestD, bestI, k)
    return bestD, bestI


class IndexFlatIP:
    """The slice of ``faiss.IndexFlatIP`` the reference touches (VectorStore_Faiss.py:115,202,263,381)."""

    metric = "ip"

    def __init__(self, d: int):
        self.d = int(d)
        self.is_trained = True
        self._x = np.zeros((0, self.d), np.float32)

    @property
    def ntotal(self) -> int:
        return self._x.shape[0]

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._x = np.concatenate([self._x, x], axis=0)

    def reset(self) -> None:
        self._x = np.zeros((0, self.d), np.float32)

    def search(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        # large blocks: FAISS scans the whole matrix per query; small blocks would only add Python overhead
        return flat_ip_search(self._x, q, k, block=1 << 20)


class IndexFlatL2(IndexFlatIP):
    """``faiss.IndexFlatL2`` (VectorStore_Faiss.py:126): squared L2 distance, ascending."""

    metric = "l2"

    def search(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        x = self._x
        d2 = ((q * q).sum(1)[:, None] - 2.0 * (q @ x.T) + (x * x).sum(1)[None, :]).astype(np.float32)
        D, I = topk_desc_stable(-d2, k)
        return (-D).astype(np.float32), I
