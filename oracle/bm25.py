"""Oracle (test infrastructure): BM25 scoring as the reference obtains it from ``rank_bm25``.

Call sites restated (relative to /root/reference):

* ``core/retrieval/bm25.py:213-218``  ``BM25Okapi(texts_processed, **bm25_params)``
* ``core/retrieval/bm25.py:302-311``  ``scores = vectorizer.get_scores(tokens)`` then
  ``np.argsort(scores)[::-1][:k]`` with ``k = min(k, len(docs))`` (``:297-298``)

``rank_bm25`` is a third-party package that is neither vendored under /root/reference nor
pinned (requirements.txt:1 lists only ``dill``) and cannot be installed here: **parity
unpinned** for its arithmetic.  This file restates the published 0.2.x ``BM25Okapi`` algorithm:

    avgdl      = sum(len(doc)) / N
    idf[t]     = log(N - n_t + 0.5) - log(n_t + 0.5)          (math.log, per distinct term in
                                                                first-seen order across the corpus)
    idf[t] < 0 -> epsilon * mean(idf)                          (mean over ALL raw idf values)
    score[d]  += idf[t] * ( tf * (k1 + 1) / ( tf + k1 * (1 - b + b * dl[d] / avgdl) ) )
                 for every query token t in order, duplicates repeated; unseen token -> idf 0

All fp64, every operation individually rounded (numpy elementwise, no FMA).
``BM25Okapi`` below is the faithful dict-based form (slow, O(N) Python work per query term, like
the original); ``Bm25Csr`` is a vectorised CSR form that is bit-identical to it (tested) and
fast enough to act as the checker at the 100k-document configuration.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence

import numpy as np

__all__ = ["BM25Okapi", "Bm25Csr", "argsort_topk", "stable_topk"]


class BM25Okapi:
    """Faithful restatement; attribute names follow rank_bm25 so it can stand in as a shim."""

    def __init__(self, corpus: Sequence[Sequence[str]], tokenizer=None, k1: float = 1.5,
                 b: float = 0.75, epsilon: float = 0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.tokenizer = tokenizer
        if tokenizer is not None:
            corpus = [tokenizer(doc) for doc in corpus]
        self.corpus_size = 0
        self.doc_len: List[int] = []
        self.doc_freqs: List[Dict[str, int]] = []
        self.idf: Dict[str, float] = {}
        containing: Dict[str, int] = {}
        total_len = 0
        for tokens in corpus:
            self.doc_len.append(len(tokens))
            total_len += len(tokens)
            tf: Dict[str, int] = {}
            for tok in tokens:
                tf[tok] = tf.get(tok, 0) + 1
            self.doc_freqs.append(tf)
            for tok in tf:
                containing[tok] = containing.get(tok, 0) + 1
            self.corpus_size += 1
        self.avgdl = total_len / self.corpus_size
        self.nd = containing
        idf_sum = 0.0
        floored = []
        for tok, n_t in containing.items():
            val = math.log(self.corpus_size - n_t + 0.5) - math.log(n_t + 0.5)
            self.idf[tok] = val
            idf_sum += val
            if val < 0:
                floored.append(tok)
        self.average_idf = idf_sum / len(self.idf)
        floor_value = self.epsilon * self.average_idf
        for tok in floored:
            self.idf[tok] = floor_value

    def get_scores(self, query: Iterable[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        dl = np.array(self.doc_len)
        for tok in query:
            tf = np.array([(d.get(tok) or 0) for d in self.doc_freqs])
            score += (self.idf.get(tok) or 0) * (
                tf * (self.k1 + 1) / (tf + self.k1 * (1 - self.b + self.b * dl / self.avgdl)))
        return score


class Bm25Csr:
    """Same numbers as ``BM25Okapi`` from CSR postings (term -> sorted doc ids, tf)."""

    def __init__(self, corpus: Sequence[Sequence[str]], k1: float = 1.5, b: float = 0.75,
                 epsilon: float = 0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        N = len(corpus)
        self.corpus_size = N
        vocab: Dict[str, int] = {}
        post_doc: List[List[int]] = []
        post_tf: List[List[int]] = []
        dl = np.zeros(N, np.int64)
        for di, tokens in enumerate(corpus):
            dl[di] = len(tokens)
            tf: Dict[str, int] = {}
            for tok in tokens:
                tf[tok] = tf.get(tok, 0) + 1
            for tok, c in tf.items():
                ti = vocab.get(tok)
                if ti is None:
                    ti = len(vocab)
                    vocab[tok] = ti
                    post_doc.append([])
                    post_tf.append([])
                post_doc[ti].append(di)
                post_tf[ti].append(c)
        self.vocab = vocab
        self.doc_len = dl
        self.avgdl = int(dl.sum()) / N
        V = len(vocab)
        df = np.array([len(p) for p in post_doc], np.int64)
        self.indptr = np.zeros(V + 1, np.int64)
        np.cumsum(df, out=self.indptr[1:])
        self.post_doc = np.fromiter((d for p in post_doc for d in p), np.int32, int(df.sum()))
        self.post_tf = np.fromiter((c for p in post_tf for c in p), np.int32, int(df.sum()))
        idf = np.empty(V, np.float64)
        idf_sum = 0.0
        for ti in range(V):            # first-seen order == vocab index order
            n_t = int(df[ti])
            val = math.log(N - n_t + 0.5) - math.log(n_t + 0.5)
            idf[ti] = val
            idf_sum += val
        self.average_idf = idf_sum / V
        idf[idf < 0] = self.epsilon * self.average_idf
        self.idf = idf
        # per-document length normaliser, same operation order as the reference expression
        self.doc_norm = self.k1 * (1 - self.b + self.b * dl / self.avgdl)

    def get_scores(self, query: Iterable[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        for tok in query:
            ti = self.vocab.get(tok)
            if ti is None:
                continue                       # idf 0 -> adds +0.0 everywhere
            lo, hi = self.indptr[ti], self.indptr[ti + 1]
            docs = self.post_doc[lo:hi]
            tf = self.post_tf[lo:hi].astype(np.int64)
            contrib = self.idf[ti] * (tf * (self.k1 + 1) / (tf + self.doc_norm[docs]))
            score[docs] += contrib             # doc ids are unique within one posting list
        return score


def argsort_topk(scores: np.ndarray, k: int) -> np.ndarray:
    """The reference's own idiom (core/retrieval/bm25.py:309): ``np.argsort(scores)[::-1][:k]``.
    Order inside groups of equal scores is numpy-implementation-defined."""
    return np.argsort(scores)[::-1][:k]


def stable_topk(scores: np.ndarray, k: int) -> np.ndarray:
    """Deterministic variant used for comparisons: descending score, ties by ascending id."""
    n = scores.shape[0]
    k = min(k, n)
    return np.lexsort((np.arange(n), -scores))[:k]
