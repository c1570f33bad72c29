"""Tensor-level entry points: torch CUDA tensors in, torch CUDA tensors out, every operation a
call into the C ABI (``_native.lib``).  PyTorch is used for device buffers and streams only."""
from __future__ import annotations

import ctypes
import threading
from typing import Optional, Tuple

import torch

from . import _native as N

_DT = {torch.float32: N.F32, torch.bfloat16: N.BF16, torch.float16: N.F16}
_ws_lock = threading.Lock()
_ws_cache: dict = {}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise TypeError(f"unsupported dtype {dt}; use float32, bfloat16 or float16") from None


def _cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise N.RagArcError(f"{name} must be a CUDA tensor: rag_arc_b200 has no CPU path")
    if not t.is_contiguous():
        raise N.RagArcError(f"{name} must be contiguous")
    return t


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_tls = threading.local()


class WorkspaceScope:
    """Scratch buffers owned by one user instead of the per-stream cache.  A CUDA graph bakes the raw
    pointers of the workspaces its kernels were captured with, so everything that captures (search
    pipelines, ``capture_search``) runs its warm-up AND its capture inside a scope of its own and keeps
    the scope alive next to the graph: no later call on the same (pooled) stream can grow, replace or
    share those buffers.  Buffers that had to grow inside the scope are retired, not freed."""

    def __init__(self):
        self.bufs: dict = {}
        self.retired: list = []
        self._prev = None

    def __enter__(self):
        self._prev = getattr(_tls, "scope", None)
        _tls.scope = self
        return self

    def __exit__(self, *exc):
        _tls.scope = self._prev
        return False

    def get(self, device, nbytes: int, tag: str) -> torch.Tensor:
        key = (device.index, tag)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                self.retired.append(buf)
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self.bufs[key] = buf
        return buf


def _workspace(device, nbytes: int, tag: str) -> torch.Tensor:
    """Scratch buffer: the active ``WorkspaceScope``'s if there is one (graph captures), else a
    grow-only buffer per (device, stream, purpose) - stream order makes that reuse safe."""
    scope = getattr(_tls, "scope", None)
    if scope is not None:
        return scope.get(device, nbytes, tag)
    key = (device.index, _stream_ptr(device), tag)
    with _ws_lock:
        buf = _ws_cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            _ws_cache[key] = buf
    return buf


def release_workspaces() -> None:
    with _ws_lock:
        _ws_cache.clear()


def normalize_cast(src: torch.Tensor, dtype: torch.dtype = torch.float32, normalize: bool = True,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [n,d] -> L2-normalised rows cast to ``dtype`` (faiss.normalize_L2 semantics)."""
    _cuda(src, "src")
    if src.dtype != torch.float32 or src.dim() != 2:
        raise N.RagArcError("normalize_cast expects a 2-D float32 tensor")
    n, d = src.shape
    if out is None:
        out = torch.empty((n, d), dtype=dtype, device=src.device)
    with torch.cuda.device(src.device):
        N.check(N.lib.ragarc_normalize_cast(src.data_ptr(), out.data_ptr(), n, d, dtype_code(out.dtype),
                                            int(bool(normalize)), _stream_ptr(src.device)), "normalize_cast")
    return out


def l2_aug_dim(d: int, dtype: torch.dtype) -> int:
    return int(N.lib.ragarc_l2_aug_dim(int(d), dtype_code(dtype)))


def l2_augment(src: torch.Tensor, dtype: torch.dtype, *, is_query: bool, normalize: bool = False,
               out: Optional[torch.Tensor] = None, sqnorm: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [n,d] -> [n, l2_aug_dim(d)] of ``dtype``: rows as [x | -|x|^2/2], queries as [q | 1]
    (squared-L2 search on the inner-product kernels, see ragarc_b200.h)."""
    _cuda(src, "src")
    if src.dtype != torch.float32 or src.dim() != 2:
        raise N.RagArcError("l2_augment expects a 2-D float32 tensor")
    n, d = src.shape
    da = l2_aug_dim(d, dtype)
    if out is None:
        out = torch.empty((n, da), dtype=dtype, device=src.device)
    with torch.cuda.device(src.device):
        N.check(N.lib.ragarc_l2_augment(src.data_ptr(), out.data_ptr(), n, d, dtype_code(dtype), int(bool(is_query)),
                                        int(bool(normalize)), sqnorm.data_ptr() if sqnorm is not None else None,
                                        _stream_ptr(src.device)), "l2_augment")
    return out


def l2_distances(scores: torch.Tensor, queries_aug: torch.Tensor, d: int) -> torch.Tensor:
    """In place: the kept values of a search over augmented matrices -> squared L2 distances."""
    _cuda(scores, "scores"); _cuda(queries_aug, "queries_aug")
    nq, k = scores.shape
    with torch.cuda.device(scores.device):
        N.check(N.lib.ragarc_l2_distances(scores.data_ptr(), queries_aug.data_ptr(), dtype_code(queries_aug.dtype), nq, k,
                                          int(d), _stream_ptr(scores.device)), "l2_distances")
    return scores


def dense_workspace_bytes(n: int, d: int, dtype: torch.dtype, nq: int, k: int) -> int:
    return int(N.lib.ragarc_dense_topk_workspace_bytes(n, d, dtype_code(dtype), nq, k))


def dense_topk(corpus: torch.Tensor, queries: torch.Tensor, k: int, *, n_rows: Optional[int] = None,
               path: int = N.DENSE_AUTO, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               return_path: bool = False):
    """Exact inner-product top-k of every query against rows ``[0, n_rows)`` of ``corpus``.

    Returns ``(scores float32 [nq,k] descending, ids int64 [nq,k])``; ids are -1 (scores -inf)
    past ``n_rows`` results.  Equal scores are ordered by ascending row id.
    """
    _cuda(corpus, "corpus"); _cuda(queries, "queries")
    if corpus.dtype != queries.dtype:
        raise N.RagArcError(f"corpus {corpus.dtype} and queries {queries.dtype} must share a dtype")
    if corpus.dim() != 2 or queries.dim() != 2 or corpus.shape[1] != queries.shape[1]:
        raise N.RagArcError(f"shape mismatch: corpus {tuple(corpus.shape)} queries {tuple(queries.shape)}")
    n = corpus.shape[0] if n_rows is None else int(n_rows)
    d = corpus.shape[1]
    nq = queries.shape[0]
    dev = corpus.device
    if out is None:
        scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    else:
        scores, ids = out
    code = dtype_code(corpus.dtype)
    wsb = int(N.lib.ragarc_dense_topk_workspace_bytes(n, d, code, nq, k))
    ws = _workspace(dev, wsb, "dense")
    used = ctypes.c_int(0)
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_dense_topk(corpus.data_ptr(), n, d, code, queries.data_ptr(), nq, k,
                                        scores.data_ptr(), ids.data_ptr(), ws.data_ptr(), ws.numel(),
                                        path, ctypes.byref(used), _stream_ptr(dev)), "dense_topk")
    if return_path:
        return scores, ids, used.value
    return scores, ids


def dense_topk_phase(phase: int, corpus: torch.Tensor, queries: Optional[torch.Tensor], nq: int, k: int, *,
                     n_rows: Optional[int] = None, id_base: int = 0, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                     out_keys: Optional[torch.Tensor] = None, inbox_table: Optional[torch.Tensor] = None, rank: int = 0,
                     nq_per_rank: int = 0, signal: bool = False, workspace_clean: bool = False, path: int = N.DENSE_AUTO):
    """One phase of a search (``ragarc_dense_topk_ex``): ``N.PHASE_SCORE`` fills the candidate lists in
    the active workspace, ``N.PHASE_SELECT`` merges them into ``out`` = (scores, ids) / ``out_keys`` /
    the owners' inboxes.  Both calls must see the SAME workspace: run them inside one
    ``WorkspaceScope`` (they may be on different streams; the caller orders them with events)."""
    _cuda(corpus, "corpus")
    n = corpus.shape[0] if n_rows is None else int(n_rows)
    d = corpus.shape[1]
    dev = corpus.device
    code = dtype_code(corpus.dtype)
    ws = _workspace(dev, int(N.lib.ragarc_dense_topk_workspace_bytes(n, d, code, nq, k)), "dense")
    o = N.DenseOpts()
    o.phase, o.id_base = int(phase), int(id_base)
    o.workspace_clean = int(bool(workspace_clean))
    if phase != N.PHASE_SCORE:
        if inbox_table is not None:
            o.inboxes, o.n_ranks, o.rank, o.nq_per_rank, o.signal = (inbox_table.data_ptr(), inbox_table.numel(), int(rank),
                                                                      int(nq_per_rank), int(bool(signal)))
        else:
            if out is not None:
                o.out_scores, o.out_ids = out[0].data_ptr(), out[1].data_ptr()
            if out_keys is not None:
                o.out_keys = out_keys.data_ptr()
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_dense_topk_ex(corpus.data_ptr(), n, d, code, queries.data_ptr() if queries is not None else None,
                                           nq, k, ctypes.byref(o), ws.data_ptr(), ws.numel(), path, None, _stream_ptr(dev)),
                "dense_topk_ex")


def normalize_split3(src: torch.Tensor, normalize: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [n,d] -> three bf16 planes [n,3d] (v = v1+v2+v3 to 2^-24), optionally L2-normalised first."""
    _cuda(src, "src")
    if src.dtype != torch.float32 or src.dim() != 2:
        raise N.RagArcError("normalize_split3 expects a 2-D float32 tensor")
    n, d = src.shape
    if out is None:
        out = torch.empty((n, 3 * d), dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device(src.device):
        N.check(N.lib.ragarc_normalize_split3(src.data_ptr(), out.data_ptr(), n, d, int(bool(normalize)),
                                              _stream_ptr(src.device)), "normalize_split3")
    return out


def dense_topk_x3(corpus_planes: torch.Tensor, query_planes: torch.Tensor, k: int, *,
                  n_rows: Optional[int] = None, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """fp32-accurate exact top-k on the tensor cores over bf16x3 planes ([n,3d] / [nq,3d])."""
    _cuda(corpus_planes, "corpus_planes"); _cuda(query_planes, "query_planes")
    if corpus_planes.dtype != torch.bfloat16 or query_planes.dtype != torch.bfloat16:
        raise N.RagArcError("bf16x3 planes must be bfloat16")
    n = corpus_planes.shape[0] if n_rows is None else int(n_rows)
    d3 = corpus_planes.shape[1]
    if d3 % 3 or query_planes.shape[1] != d3:
        raise N.RagArcError("plane matrices must be [*, 3d] with equal d")
    d = d3 // 3
    nq = query_planes.shape[0]
    dev = corpus_planes.device
    if out is None:
        scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    else:
        scores, ids = out
    wsb = int(N.lib.ragarc_dense_topk_x3_workspace_bytes(n, d, nq, k))
    if wsb == 0:
        raise N.RagArcError(f"bf16x3 search unsupported for d={d}, k={k}")
    ws = _workspace(dev, wsb, "dense")
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_dense_topk_x3(corpus_planes.data_ptr(), n, d, query_planes.data_ptr(), nq, k,
                                           scores.data_ptr(), ids.data_ptr(), ws.data_ptr(), ws.numel(),
                                           _stream_ptr(dev)), "dense_topk_x3")
    return scores, ids


def dense_topk_keys(corpus: torch.Tensor, queries: torch.Tensor, k: int, id_base: int, *,
                    n_rows: Optional[int] = None, path: int = N.DENSE_AUTO,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-shard search returning packed sortable keys (int64 view of uint64) for the all-gather."""
    _cuda(corpus, "corpus"); _cuda(queries, "queries")
    n = corpus.shape[0] if n_rows is None else int(n_rows)
    d = corpus.shape[1]
    nq = queries.shape[0]
    dev = corpus.device
    keys = out if out is not None else torch.empty((nq, k), dtype=torch.int64, device=dev)
    code = dtype_code(corpus.dtype)
    ws = _workspace(dev, int(N.lib.ragarc_dense_topk_workspace_bytes(n, d, code, nq, k)), "dense")
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_dense_topk_keys(corpus.data_ptr(), n, d, code, queries.data_ptr(), nq, k,
                                             int(id_base), keys.data_ptr(), ws.data_ptr(), ws.numel(),
                                             path, None, _stream_ptr(dev)), "dense_topk_keys")
    return keys


def dense_topk_keys_push(corpus: torch.Tensor, queries: torch.Tensor, k: int, id_base: int,
                         inbox_table: torch.Tensor, rank: int, nq_per_rank: int, *, signal: bool = False,
                         n_rows: Optional[int] = None, path: int = N.DENSE_AUTO) -> None:
    """Per-shard search whose sorted key rows go straight into the inbox of the rank owning each
    query (``inbox_table``: int64 device tensor of one ``[n_ranks, nq_per_rank, k]`` inbox pointer per
    rank, local or peer memory) - the producer half of the query-owner exchange.  ``signal``: also
    bump the owner's per-query arrival counters behind the inbox (see ``merge_topk_inbox``)."""
    _cuda(corpus, "corpus"); _cuda(queries, "queries"); _cuda(inbox_table, "inbox_table")
    n = corpus.shape[0] if n_rows is None else int(n_rows)
    d = corpus.shape[1]
    nq = queries.shape[0]
    dev = corpus.device
    code = dtype_code(corpus.dtype)
    ws = _workspace(dev, int(N.lib.ragarc_dense_topk_workspace_bytes(n, d, code, nq, k)), "dense")
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_dense_topk_keys_push(corpus.data_ptr(), n, d, code, queries.data_ptr(), nq, k,
                                                  int(id_base), inbox_table.data_ptr(), inbox_table.numel(),
                                                  int(rank), int(nq_per_rank), int(bool(signal)), ws.data_ptr(),
                                                  ws.numel(), path, None, _stream_ptr(dev)), "dense_topk_keys_push")


def inbox_words(n_ranks: int, nq_per_rank: int, k: int) -> int:
    """int64 words of one signalled inbox: the key block + one uint32 arrival counter per owned query."""
    return n_ranks * nq_per_rank * k + (nq_per_rank + 1) // 2


def merge_topk_inbox(inbox: torch.Tensor, n_ranks: int, nq_per_rank: int, nq_own: int, k_in: int, k_out: int,
                     status: Optional[torch.Tensor] = None, timeout_ms: float = 2000.0):
    """Owner half of the signalled exchange: waits (inside the kernel, per query) until all ``n_ranks``
    rows of a query have arrived in ``inbox`` (int64 ``[inbox_words]``), then merges them; ``nq_own``
    of the ``nq_per_rank`` inbox rows belong to queries of this batch.  -> ``[nq_own, k_out]`` results."""
    _cuda(inbox, "inbox")
    dev = inbox.device
    scores = torch.empty((nq_own, k_out), dtype=torch.float32, device=dev)
    ids = torch.empty((nq_own, k_out), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_merge_topk_inbox(inbox.data_ptr(), n_ranks, nq_per_rank, nq_own, k_in, k_out, scores.data_ptr(),
                                              ids.data_ptr(), float(timeout_ms),
                                              status.data_ptr() if status is not None else None, _stream_ptr(dev)),
                "merge_topk_inbox")
    return scores, ids


def merge_topk_keys(keys: torch.Tensor, k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """keys: [G, nq, k_in] packed keys (as all-gathered) -> merged (scores, ids)."""
    _cuda(keys, "keys")
    G, nq, k_in = keys.shape
    dev = keys.device
    scores = torch.empty((nq, k_out), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k_out), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_merge_topk_keys(keys.data_ptr(), G, nq, k_in, k_out, scores.data_ptr(),
                                             ids.data_ptr(), _stream_ptr(dev)), "merge_topk_keys")
    return scores, ids


def merge_topk_keys_p2p(ptr_table: torch.Tensor, nq: int, k_in: int, k_out: int,
                        out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """ptr_table: int64 device tensor of G pointers to [nq,k_in] key blocks (local or peer memory)."""
    _cuda(ptr_table, "ptr_table")
    dev = ptr_table.device
    if out is None:
        scores = torch.empty((nq, k_out), dtype=torch.float32, device=dev)
        ids = torch.empty((nq, k_out), dtype=torch.int64, device=dev)
    else:
        scores, ids = out
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_merge_topk_keys_p2p(ptr_table.data_ptr(), ptr_table.numel(), nq, k_in, k_out,
                                                 scores.data_ptr(), ids.data_ptr(), _stream_ptr(dev)),
                "merge_topk_keys_p2p")
    return scores, ids


def _post_val_ptr(index, use_post_val: bool) -> int:
    pv = getattr(index, "post_val", None) if use_post_val else None
    return pv.data_ptr() if pv is not None else 0


def bm25_topk(index, q_terms: torch.Tensor, q_len: torch.Tensor, k: int, use_post_val: bool = True):
    """index: object with device tensors indptr/post_doc/post_tf/idf/doc_norm (and optionally the
    precomputed per-posting factor post_val) and k1_plus_1, n_docs."""
    _cuda(q_terms, "q_terms"); _cuda(q_len, "q_len")
    nq, tmax = q_terms.shape
    dev = q_terms.device
    scores = torch.empty((nq, k), dtype=torch.float64, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    ws = _workspace(dev, int(N.lib.ragarc_bm25_workspace_bytes(index.n_docs, max(nq, 1))), "bm25")
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_bm25_topk(index.indptr.data_ptr(), index.post_doc.data_ptr(),
                                       index.post_tf.data_ptr(), _post_val_ptr(index, use_post_val), index.idf.data_ptr(),
                                       index.doc_norm.data_ptr(), float(index.k1_plus_1),
                                       q_terms.data_ptr(), q_len.data_ptr(), nq, tmax, index.n_docs, k,
                                       scores.data_ptr(), ids.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _stream_ptr(dev)), "bm25_topk")
    base = int(getattr(index, "id_base", 0))
    if base:
        ids = torch.where(ids >= 0, ids + base, ids)     # shard-local doc ids -> global
    return scores, ids


def bm25_merge_topk(scores: torch.Tensor, ids: torch.Tensor, k_out: int):
    """scores float64 / ids int64 ``[n_lists, nq, k_in]`` (per-shard BM25 results carrying global doc
    ids, -1 padding) -> global top-k_out per query (scores float64 [nq,k_out], ids int64)."""
    _cuda(scores, "scores"); _cuda(ids, "ids")
    if scores.dtype != torch.float64 or ids.dtype != torch.int64 or scores.shape != ids.shape or scores.dim() != 3:
        raise N.RagArcError("bm25_merge_topk expects float64 scores and int64 ids of shape [n_lists, nq, k_in]")
    G, nq, k_in = scores.shape
    dev = scores.device
    out_s = torch.empty((nq, k_out), dtype=torch.float64, device=dev)
    out_i = torch.empty((nq, k_out), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_bm25_merge_topk(scores.data_ptr(), ids.data_ptr(), G, nq, k_in, k_out,
                                             out_s.data_ptr(), out_i.data_ptr(), _stream_ptr(dev)), "bm25_merge_topk")
    return out_s, out_i


def bm25_scores(index, q_terms: torch.Tensor, q_len: torch.Tensor, use_post_val: bool = True) -> torch.Tensor:
    _cuda(q_terms, "q_terms"); _cuda(q_len, "q_len")
    nq, tmax = q_terms.shape
    dev = q_terms.device
    out = torch.empty((nq, index.n_docs), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_bm25_scores(index.indptr.data_ptr(), index.post_doc.data_ptr(),
                                         index.post_tf.data_ptr(), _post_val_ptr(index, use_post_val), index.idf.data_ptr(),
                                         index.doc_norm.data_ptr(), float(index.k1_plus_1),
                                         q_terms.data_ptr(), q_len.data_ptr(), nq, tmax, index.n_docs,
                                         out.data_ptr(), _stream_ptr(dev)), "bm25_scores")
    return out


def rrf_fuse(ids: torch.Tensor, top_k: int, rrf_k: float = 60.0):
    """ids: int32 [L, nq, kl] document keys (negative = padding).  Returns (ids int32 [nq,top_k],
    scores float64 [nq,top_k], count int32 [nq])."""
    _cuda(ids, "ids")
    if ids.dtype != torch.int32 or ids.dim() != 3:
        raise N.RagArcError("rrf_fuse expects int32 [L, nq, kl]")
    L, nq, kl = ids.shape
    dev = ids.device
    out_ids = torch.empty((nq, top_k), dtype=torch.int32, device=dev)
    out_scores = torch.empty((nq, top_k), dtype=torch.float64, device=dev)
    out_count = torch.empty((nq,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_rrf_fuse(ids.data_ptr(), L, nq, kl, float(rrf_k), top_k, out_ids.data_ptr(),
                                      out_scores.data_ptr(), out_count.data_ptr(), _stream_ptr(dev)),
                "rrf_fuse")
    return out_ids, out_scores, out_count


def rrf_fuse_rows(rows, row_to_key, kl: int, top_k: int, rrf_k: float = 60.0):
    """The hybrid merge on retriever rows in one launch (ragarc_rrf_fuse_rows).  rows: one int64 [nq, <=kl]
    CUDA tensor per retriever (None = that retriever returned nothing); row_to_key: one int32 CUDA tensor
    per retriever (row -> content key).  Returns (keys int32 [nq,top_k],
    scores float64 [nq,top_k], packed) where ``packed`` is ONE uint8 CUDA buffer holding the row of every
    fused entry in its list's corpus, the list, and the number of entries per query (unpack_fused_rows) -
    one device->host transfer hands the caller everything it needs to pick the Documents."""
    L = len(rows)
    live = [r for r in rows if r is not None]
    if not live:
        raise N.RagArcError("rrf_fuse_rows: every list is empty")
    dev, nq = live[0].device, int(live[0].shape[0])
    hold = []
    rp, tp = (ctypes.c_void_p * L)(), (ctypes.c_void_p * L)()
    ke = (ctypes.c_int * L)()
    for l, r in enumerate(rows):
        t = row_to_key[l]
        _cuda(t, "row_to_key")
        if t.dtype != torch.int32:
            raise N.RagArcError("rrf_fuse_rows: row_to_key tables are int32")
        tp[l] = t.data_ptr()
        if r is None:
            rp[l], ke[l] = None, 0
            continue
        _cuda(r, "rows")
        if r.dtype != torch.int64 or r.dim() != 2 or r.shape[0] != nq or r.shape[1] > kl:
            raise N.RagArcError("rrf_fuse_rows expects int64 [nq, <=kl] rows per list")
        r = r.contiguous(); hold.append(r)
        rp[l], ke[l] = r.data_ptr(), int(r.shape[1])
    out_ids = torch.empty((nq, top_k), dtype=torch.int32, device=dev)
    out_scores = torch.empty((nq, top_k), dtype=torch.float64, device=dev)
    # one buffer, three views: [rows int64 nq*top_k | lists int32 nq*top_k | counts int32 nq]
    n = nq * top_k
    buf = torch.empty((n * 12 + nq * 4,), dtype=torch.uint8, device=dev)
    out_row = buf[:n * 8].view(torch.int64)
    out_list = buf[n * 8:n * 12].view(torch.int32)
    out_count = buf[n * 12:].view(torch.int32)
    with torch.cuda.device(dev):
        N.check(N.lib.ragarc_rrf_fuse_rows(rp, ke, tp, L, nq, kl,
                                           float(rrf_k), top_k, out_ids.data_ptr(), out_scores.data_ptr(),
                                           out_count.data_ptr(), out_list.data_ptr(), out_row.data_ptr(),
                                           _stream_ptr(dev)), "rrf_fuse_rows")
    return out_ids, out_scores, buf


def unpack_fused_rows(buf_host, nq: int, top_k: int):
    """Host views of the buffer rrf_fuse_rows returns (after .cpu().numpy()): (lists [nq,top_k] int32,
    rows [nq,top_k] int64, counts [nq] int32)."""
    n = nq * top_k
    return (buf_host[n * 8:n * 12].view("<i4").reshape(nq, top_k), buf_host[:n * 8].view("<i8").reshape(nq, top_k),
            buf_host[n * 12:].view("<i4"))


_POOL = {"mean": N.POOL_MEAN, "cls": N.POOL_CLS, "last": N.POOL_LAST}


def pool_normalize(x: torch.Tensor, mask: torch.Tensor, mode: str = "mean", normalize: bool = True
                   ) -> torch.Tensor:
    """x: [B,T,H] encoder output, mask: [B,T] (any integer/bool dtype) -> fp32 [B,H]."""
    _cuda(x, "x")
    if mode not in _POOL:
        raise ValueError(f"pooling mode must be one of {sorted(_POOL)}")
    B, T, H = x.shape
    m = mask.to(device=x.device, dtype=torch.int32).contiguous()
    out = torch.empty((B, H), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib.ragarc_pool_normalize(x.data_ptr(), dtype_code(x.dtype), m.data_ptr(), B, T, H,
                                            _POOL[mode], int(bool(normalize)), out.data_ptr(),
                                            _stream_ptr(x.device)), "pool_normalize")
    return out


def mmr_select(corpus: torch.Tensor, queries: torch.Tensor, cand_rows: torch.Tensor, k: int,
               lambda_mult: float = 0.5, *, n_rows: Optional[int] = None) -> torch.Tensor:
    """Greedy MMR over candidate rows [nq, fetch_k] (int64, -1 padded) -> int32 [nq,k] indices into
    the candidate lists."""
    _cuda(corpus, "corpus"); _cuda(queries, "queries"); _cuda(cand_rows, "cand_rows")
    n = corpus.shape[0] if n_rows is None else int(n_rows)
    nq, fetch_k = cand_rows.shape
    out = torch.empty((nq, k), dtype=torch.int32, device=corpus.device)
    with torch.cuda.device(corpus.device):
        N.check(N.lib.ragarc_mmr_select(corpus.data_ptr(), n, corpus.shape[1], dtype_code(corpus.dtype),
                                        queries.data_ptr(), nq, cand_rows.data_ptr(), fetch_k, k,
                                        float(lambda_mult), out.data_ptr(), _stream_ptr(corpus.device)),
                "mmr_select")
    return out


def adjacent_cosine_distance(x: torch.Tensor) -> torch.Tensor:
    """x: float32 or float64 ``[n,d]`` -> float64 ``[max(n-1,0)]`` cosine distances of consecutive rows
    (spliter.py:354-372)."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise N.RagArcError("x must be a CUDA tensor: rag_arc_b200 has no CPU path")
    if x.dim() != 2 or x.dtype not in (torch.float32, torch.float64):
        raise N.RagArcError("adjacent_cosine_distance expects a float32/float64 [n,d] tensor")
    x = x.contiguous()
    n, d = x.shape
    out = torch.empty((max(n - 1, 0),), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib.ragarc_adjacent_cosine_distance(x.data_ptr(), N.F32 if x.dtype == torch.float32 else N.F64, n, d,
                                                      out.data_ptr(), _stream_ptr(x.device)), "adjacent_cosine_distance")
    return out


def yes_no_score(last_logits: torch.Tensor, true_id: int, false_id: int) -> torch.Tensor:
    """last_logits: ``[B, vocab]`` (rows may be strided, e.g. ``logits[:, -1, :]``) -> float32 ``[B]``
    P(yes) as Reranker_Qwen3.py:44-49 computes it."""
    if not isinstance(last_logits, torch.Tensor) or not last_logits.is_cuda:
        raise N.RagArcError("last_logits must be a CUDA tensor: rag_arc_b200 has no CPU path")
    if last_logits.dim() != 2 or last_logits.stride(1) != 1:
        raise N.RagArcError("yes_no_score expects [B, vocab] with unit stride along the vocabulary")
    B, V = last_logits.shape
    out = torch.empty((B,), dtype=torch.float32, device=last_logits.device)
    with torch.cuda.device(last_logits.device):
        N.check(N.lib.ragarc_yes_no_score(last_logits.data_ptr(), dtype_code(last_logits.dtype), B,
                                          last_logits.stride(0) if B > 1 else V, V, int(true_id), int(false_id),
                                          out.data_ptr(), _stream_ptr(last_logits.device)), "yes_no_score")
    return out
