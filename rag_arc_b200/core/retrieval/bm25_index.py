"""Host-side BM25 index build: tokenised corpus -> CSR postings + fp64 tables on the GPU.

Builds exactly the statistics ``rank_bm25.BM25Okapi.__init__`` derives (the constructor the
reference calls at /root/reference core/retrieval/bm25.py:218): per-document term frequencies and
lengths, ``avgdl``, document frequencies, ``idf = log(N-n+0.5) - log(n+0.5)`` with the
``epsilon * mean(idf)`` floor for negative values, defaults ``k1=1.5, b=0.75, epsilon=0.25``.
The idf table is computed with Python ``math.log`` and summed in first-seen vocabulary order so
that every fp64 value is bit-identical to what the reference holds; the per-document length
normaliser ``k1*(1-b+b*dl/avgdl)`` is evaluated with numpy in the reference's operation order.
Index build is host work (not the optimisation target); scoring is ``ragarc_bm25_topk``.
"""
from __future__ import annotations

import math
import threading
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch


class _VocabHandle:
    """Owns a ``ragarc_vocab_t`` (freed with the index)."""

    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            from ... import _native as N
            N.lib.ragarc_vocab_free(self.h)
        except Exception:
            pass


class Bm25Index:
    def __init__(self, *, vocab: Optional[Dict[str, int]], indptr: np.ndarray, post_doc: np.ndarray,
                 post_tf: np.ndarray, doc_len: np.ndarray, k1: float, b: float, epsilon: float,
                 device):
        self.vocab = vocab
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.n_docs = int(doc_len.shape[0])
        self.doc_len_np = doc_len.astype(np.int64)
        self.avgdl = int(self.doc_len_np.sum()) / self.n_docs
        df = np.diff(indptr)
        V = int(df.shape[0])
        N = self.n_docs
        idf = np.empty(V, np.float64)
        total = 0.0
        log = math.log
        for ti in range(V):          # first-seen order, sequential Python-float sum
            n_t = int(df[ti])
            val = log(N - n_t + 0.5) - log(n_t + 0.5)
            idf[ti] = val
            total += val
        self.average_idf = total / V if V else 0.0
        idf[idf < 0] = epsilon * self.average_idf
        self.idf_np = idf
        self.doc_norm_np = k1 * (1 - b + b * self.doc_len_np / self.avgdl)
        self.k1_plus_1 = k1 + 1
        self.indptr_np, self.post_doc_np, self.post_tf_np = indptr, post_doc, post_tf
        # query-independent factor of every posting, evaluated with numpy in the reference's order:
        # tf * (k1 + 1) / (tf + k1 * (1 - b + b * dl / avgdl))
        tf64 = post_tf.astype(np.int64)
        self.post_val_np = tf64 * (k1 + 1) / (tf64 + self.doc_norm_np[post_doc])
        self.id_base = 0
        self.device = None
        if device is not None:
            self._upload(device)

    def _upload(self, device):
        self.device = torch.device(device)
        self.indptr = torch.from_numpy(self.indptr_np.astype(np.int64)).to(self.device)
        self.post_doc = torch.from_numpy(self.post_doc_np.astype(np.int32)).to(self.device)
        self.post_tf = torch.from_numpy(self.post_tf_np.astype(np.int32)).to(self.device)
        self.post_val = torch.from_numpy(self.post_val_np).to(self.device)
        self.idf = torch.from_numpy(self.idf_np).to(self.device)
        self.doc_norm = torch.from_numpy(self.doc_norm_np).to(self.device)

    def shard(self, lo: int, hi: int, device) -> "Bm25Index":
        """The index of documents ``[lo, hi)`` for doc-range sharded (multi-GPU) search: postings and
        length normalisers of those documents only, local doc ids ``doc - lo`` (``id_base = lo`` maps
        results back), but the GLOBAL idf table, average length and vocabulary - the reference
        derives all three from the whole corpus (bm25.py:218), so every shard must use the same
        values for the merged result to equal the single-index one bit for bit."""
        import copy
        sh = copy.copy(self)
        keep = (self.post_doc_np >= lo) & (self.post_doc_np < hi)
        V = len(self.indptr_np) - 1
        term_of = np.repeat(np.arange(V, dtype=np.int64), np.diff(self.indptr_np))
        sh.indptr_np = np.zeros(V + 1, np.int64)
        np.cumsum(np.bincount(term_of[keep], minlength=V), out=sh.indptr_np[1:])
        sh.post_doc_np = (self.post_doc_np[keep] - lo).astype(np.int32)
        sh.post_tf_np = self.post_tf_np[keep]
        sh.post_val_np = self.post_val_np[keep]
        sh.doc_len_np = self.doc_len_np[lo:hi]
        sh.doc_norm_np = self.doc_norm_np[lo:hi]
        sh.n_docs = int(hi - lo)
        sh.id_base = int(lo)
        sh._upload(device)
        return sh

    # -- construction --------------------------------------------------------------------------
    @classmethod
    def from_token_lists(cls, corpus: Sequence[Sequence[str]], *, k1: float = 1.5, b: float = 0.75,
                         epsilon: float = 0.25, device="cuda") -> "Bm25Index":
        vocab: Dict[str, int] = {}
        flat: List[int] = []
        lens = np.empty(len(corpus), np.int64)
        get = vocab.get
        for di, tokens in enumerate(corpus):
            lens[di] = len(tokens)
            for tok in tokens:
                ti = get(tok)
                if ti is None:
                    ti = len(vocab)
                    vocab[tok] = ti
                flat.append(ti)
        toks = np.asarray(flat, dtype=np.int64)
        offs = np.zeros(len(corpus) + 1, np.int64)
        np.cumsum(lens, out=offs[1:])
        return cls._from_ids(toks, offs, len(vocab), vocab, k1, b, epsilon, device)

    @classmethod
    def from_token_ids(cls, toks: np.ndarray, offs: np.ndarray, *, k1: float = 1.5, b: float = 0.75,
                       epsilon: float = 0.25, device="cuda") -> "Bm25Index":
        """Corpus given as integer token ids (synthetic corpora).  Term ids are re-numbered in
        first-seen order, which is the order the reference's idf mean is accumulated in; the
        vocabulary maps the *original* id (as ``t<id>`` string and as int) to the new one."""
        toks = np.asarray(toks, dtype=np.int64)
        uniq, first = np.unique(toks, return_index=True)
        order = np.argsort(first, kind="stable")
        remap = np.empty(int(uniq.max()) + 1, np.int64)
        remap[:] = -1
        remap[uniq[order]] = np.arange(len(uniq))
        idx = cls._from_ids(remap[toks], np.asarray(offs, np.int64), len(uniq), None, k1, b, epsilon, device)
        idx.id_remap = remap
        return idx

    @classmethod
    def _from_ids(cls, toks, offs, V, vocab, k1, b, epsilon, device):
        n_docs = len(offs) - 1
        lens = np.diff(offs)
        doc_of = np.repeat(np.arange(n_docs, dtype=np.int64), lens)
        pair = toks * n_docs + doc_of                 # sorts by term, then doc
        upair, tf = np.unique(pair, return_counts=True)
        term = upair // n_docs
        post_doc = (upair % n_docs).astype(np.int32)
        indptr = np.zeros(V + 1, np.int64)
        np.cumsum(np.bincount(term, minlength=V), out=indptr[1:])
        return cls(vocab=vocab, indptr=indptr, post_doc=post_doc, post_tf=tf.astype(np.int32),
                   doc_len=lens, k1=k1, b=b, epsilon=epsilon, device=device)

    # -- queries -------------------------------------------------------------------------------
    def _native_vocab(self):
        """The vocabulary as a C++ hash table inside the library (built once, on first use)."""
        v = getattr(self, "_vocab_handle", None)
        if v is None:
            import ctypes
            from ... import _native as N
            toks = [None] * len(self.vocab)
            for tok, ti in self.vocab.items():
                toks[ti] = tok.encode("utf-8")
            offs = np.zeros(len(toks) + 1, np.int64)
            np.cumsum([len(t) for t in toks], out=offs[1:])
            blob = b"".join(toks)
            h = ctypes.c_void_p()
            N.check(N.lib.ragarc_vocab_create(blob, offs.ctypes.data, len(toks), ctypes.byref(h)), "vocab_create")
            v = self._vocab_handle = _VocabHandle(h)
        return v.h

    def encode_texts(self, texts: Sequence[str], tmax: int = 32):
        """Whitespace-tokenise (``str.split()`` semantics), look up and pack a batch of query strings in
        one library call (``ragarc_vocab_encode_split``) -> the ``(q_terms, q_len)`` device tensors
        ``encode_queries`` returns for ``[t.split() for t in texts]``.  The packed arrays are written
        straight into pinned staging buffers and uploaded asynchronously."""
        import ctypes
        from ... import _native as N
        h = self._native_vocab()
        nq = len(texts)
        joined = "\0".join(texts)
        nul_ok = joined.count("\0") == max(nq - 1, 0)           # no query contains a NUL itself (else: offsets form)
        if nul_ok:
            blob, offs = joined.encode("utf-8"), None
        else:
            enc = [t.encode("utf-8") for t in texts]
            offs = np.zeros(nq + 1, np.int64)
            np.cumsum([len(b) for b in enc], out=offs[1:])
            blob = b"".join(enc)
        cuda = self.device is not None and self.device.type == "cuda"
        lock = self.__dict__.setdefault("_enc_lock", threading.Lock())       # retrievers run from thread pools
        with lock:
            while True:
                stage = getattr(self, "_enc_stage", None)
                if stage is None or stage[0].shape[0] < nq or stage[0].shape[1] != tmax:
                    terms = torch.empty((max(nq, 1), tmax), dtype=torch.int32)
                    lens = torch.empty((max(nq, 1),), dtype=torch.int32)
                    if cuda:
                        terms, lens = terms.pin_memory(), lens.pin_memory()
                    stage = self._enc_stage = (terms, lens, torch.cuda.Event() if cuda else None)
                elif stage[2] is not None:
                    stage[2].synchronize()          # the previous asynchronous upload out of the buffers is done
                longest = ctypes.c_int(0)
                if nul_ok:
                    N.check(N.lib.ragarc_vocab_encode_split0(h, blob, len(blob), nq, tmax, stage[0].data_ptr(),
                                                             stage[1].data_ptr(), ctypes.byref(longest)), "vocab_encode_split0")
                else:
                    N.check(N.lib.ragarc_vocab_encode_split(h, blob, offs.ctypes.data, nq, tmax, stage[0].data_ptr(),
                                                            stage[1].data_ptr(), ctypes.byref(longest)), "vocab_encode_split")
                if longest.value <= tmax:
                    break
                tmax = 1 << (longest.value - 1).bit_length()              # a longer query than the staging row: grow, redo
            width = max(1, longest.value)
            q_terms = stage[0][:nq, :width].to(self.device, non_blocking=True).contiguous()
            q_len = stage[1][:nq].to(self.device, non_blocking=True)
            if stage[2] is not None:
                stage[2].record(torch.cuda.current_stream(self.device))
        return q_terms, q_len

    def encode_queries(self, queries: Iterable[Sequence[str]]):
        """Token lists -> (q_terms int32 [nq,tmax] with -1 for out-of-vocabulary tokens, q_len)."""
        get = self.vocab.get
        queries = [q if isinstance(q, (list, tuple)) else list(q) for q in queries]
        lens = np.fromiter((len(q) for q in queries), np.int32, len(queries))
        flat = np.fromiter((get(tok, -1) for q in queries for tok in q), np.int32, int(lens.sum()))
        return self._pack_flat(flat, lens)

    def encode_query_ids(self, qids: np.ndarray):
        """Integer-token queries for an index built with ``from_token_ids``."""
        qids = np.asarray(qids, np.int64)
        safe = np.where((qids >= 0) & (qids < len(self.id_remap)), qids, 0)
        mapped = np.where((qids >= 0) & (qids < len(self.id_remap)), self.id_remap[safe], -1)
        q_terms = torch.from_numpy(mapped.astype(np.int32)).to(self.device)
        q_len = torch.full((qids.shape[0],), qids.shape[1], dtype=torch.int32, device=self.device)
        return q_terms.contiguous(), q_len

    def _pack_flat(self, flat: np.ndarray, lens: np.ndarray):
        """Ragged rows given as one flat id array + row lengths -> padded [nq, tmax] (-1) + lengths."""
        tmax = max(1, int(lens.max()) if lens.size else 1)
        arr = np.full((lens.shape[0], tmax), -1, np.int32)
        arr[np.arange(tmax, dtype=np.int32)[None, :] < lens[:, None]] = flat       # row-major fill
        return (torch.from_numpy(arr).to(self.device), torch.from_numpy(lens.astype(np.int32)).to(self.device))

    def _pack(self, rows):
        lens = np.fromiter((len(r) for r in rows), np.int32, len(rows))
        flat = np.fromiter((t for r in rows for t in r), np.int32, int(lens.sum()))
        return self._pack_flat(flat, lens)
