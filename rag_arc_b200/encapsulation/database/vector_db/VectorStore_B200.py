"""``B200VectorStore``: drop-in for ``FaissVectorStore`` (/root/reference
encapsulation/database/vector_db/VectorStore_Faiss.py) whose index lives in B200 HBM.

Same constructor arguments, attributes (``embedding``, ``index``, ``index_type``, ``metric``,
``normalize_L2``, ``docstore``, ``index_to_docstore_id``), methods, return types and errors as the
reference class (:68-513).  What changes underneath:

* the corpus is one contiguous row-major ``[capacity, d]`` device matrix (fp32 by default - the
  reference's precision - or bf16/fp16 via ``dtype=`` for the tensor-core path), rows appended at
  ``ntotal`` exactly like ``IndexFlatIP.add`` (:199-208);
* ``faiss.normalize_L2`` (:150-154) and the fp32 staging casts (:170,:258) are
  ``ragarc_normalize_cast``; ``IndexFlatIP.search`` (:263) is ``ragarc_dense_topk`` (scoring fused
  with per-query top-k);
* MMR gathers candidate rows from the resident matrix instead of re-embedding every candidate
  (:301-304) and runs ``ragarc_mmr_select``;
* ``delete`` compacts rows on the device instead of re-embedding the survivors (:374-415);
* batched entry points (``search_batch``, ``similarity_search_batch``) sit next to the
  single-query methods - the reference has none.

Only exact ("flat") search is offered - all three metrics of the reference (cosine, ip, l2);
``index_type`` "ivf"/"hnsw" (approximate in the reference) raise ``ValueError``.
"""
from __future__ import annotations

import os
import pickle
import threading
import uuid
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ....core.utils.data_model import Document
from .... import ops
from .VectorStoreBase import VectorStore

_DTYPES = {"float32": torch.float32, "fp32": torch.float32, "bfloat16": torch.bfloat16, "bf16": torch.bfloat16,
           "float16": torch.float16, "fp16": torch.float16}


def _is_x3(dt) -> bool:
    return isinstance(dt, str) and dt.lower() in ("float32x3", "fp32x3", "bf16x3")


def _is_exact_fp32(dt) -> bool:
    return isinstance(dt, str) and dt.lower() in ("float32_exact", "fp32_exact")


def _as_dtype(dt) -> torch.dtype:
    if _is_x3(dt) or _is_exact_fp32(dt):
        return torch.float32
    if isinstance(dt, torch.dtype):
        if dt not in ops._DT:
            raise ValueError(f"unsupported storage dtype: {dt}")
        return dt
    try:
        return _DTYPES[str(dt).lower()]
    except KeyError:
        raise ValueError(f"unsupported storage dtype: {dt}") from None


class FlatIndexB200:
    """The slice of the ``faiss.IndexFlatIP`` API the reference uses (``d``, ``ntotal``,
    ``is_trained``, ``add``, ``reset``, ``search``) over a device-resident matrix."""

    def __init__(self, d: int, dtype: torch.dtype, device, normalize: bool, x3: bool = False, l2: bool = False):
        self.d = int(d)
        self.dtype = dtype
        self.device = torch.device(device)
        self.normalize = normalize
        self.is_trained = True
        self.ntotal = 0
        self.rows = torch.empty((0, self.d), dtype=dtype, device=self.device)
        # metric "l2" (faiss.IndexFlatL2): next to the plain rows, the augmented matrix [x | -|x|^2/2]
        # the inner-product kernels search (ragarc_l2_augment); results are squared distances, ascending
        self._stage_lock = threading.Lock()
        self.l2 = bool(l2)
        self.d_aug = ops.l2_aug_dim(self.d, dtype) if l2 else self.d
        self.aug = torch.empty((0, self.d_aug), dtype=dtype, device=self.device) if l2 else None
        # "float32x3": fp32 master rows + three bf16 planes per row for fp32-accurate tensor-core search
        self.x3 = bool(x3)
        self.planes = torch.empty((0, 3 * self.d), dtype=torch.bfloat16, device=self.device) if x3 else None

    def _reserve(self, need: int) -> None:
        cap = self.rows.shape[0]
        if need <= cap:
            return
        new_cap = max(need, int(cap * 1.5) + 1024)
        grown = torch.empty((new_cap, self.d), dtype=self.dtype, device=self.device)
        if self.ntotal:
            grown[:self.ntotal].copy_(self.rows[:self.ntotal])
        self.rows = grown
        if self.x3:
            gp = torch.empty((new_cap, 3 * self.d), dtype=torch.bfloat16, device=self.device)
            if self.ntotal:
                gp[:self.ntotal].copy_(self.planes[:self.ntotal])
            self.planes = gp
        if self.l2:
            ga = torch.empty((new_cap, self.d_aug), dtype=self.dtype, device=self.device)
            if self.ntotal:
                ga[:self.ntotal].copy_(self.aug[:self.ntotal])
            self.aug = ga

    def add(self, x) -> None:
        """x: fp32 [n,d] numpy array or tensor (host or device); normalised (if the store's metric
        asks for it) and cast into the matrix on the device."""
        xt = torch.as_tensor(x)
        if xt.dim() != 2 or xt.shape[1] != self.d:
            raise ValueError(f"expected [n,{self.d}] vectors, got {tuple(xt.shape)}")
        xt = xt.to(device=self.device, dtype=torch.float32).contiguous()
        n = xt.shape[0]
        self._reserve(self.ntotal + n)
        ops.normalize_cast(xt, self.dtype, self.normalize, out=self.rows[self.ntotal:self.ntotal + n])
        if self.x3:
            ops.normalize_split3(xt, self.normalize, out=self.planes[self.ntotal:self.ntotal + n])
        if self.l2:
            ops.l2_augment(xt, self.dtype, is_query=False, normalize=self.normalize, out=self.aug[self.ntotal:self.ntotal + n])
        self.ntotal += n

    def add_prepared(self, rows: torch.Tensor) -> None:
        """Append rows that are already normalised and in the storage dtype (device tensor)."""
        n = rows.shape[0]
        self._reserve(self.ntotal + n)
        self.rows[self.ntotal:self.ntotal + n].copy_(rows)
        if self.x3:
            ops.normalize_split3(rows.to(torch.float32).contiguous(), False,
                                 out=self.planes[self.ntotal:self.ntotal + n])
        if self.l2:
            ops.l2_augment(rows.to(torch.float32).contiguous(), self.dtype, is_query=False, normalize=False,
                           out=self.aug[self.ntotal:self.ntotal + n])
        self.ntotal += n

    def reset(self) -> None:
        self.ntotal = 0

    def _to_device_f32(self, qt: torch.Tensor) -> torch.Tensor:
        """Host queries -> fp32 on the device.  bf16 / fp16 host tensors cross PCIe as they are (half the bytes
        of fp32) and are widened on the device, which is exact.  Pageable memory goes through a page-locked
        staging buffer owned by the index (one CPU copy + an asynchronous DMA instead of the driver's
        synchronous bounce); the buffer is only rewritten after the previous transfer out of it has completed."""
        if qt.is_cuda or self.device.type != "cuda":
            return qt.to(device=self.device, dtype=torch.float32).contiguous()
        if qt.dtype not in (torch.bfloat16, torch.float16):
            qt = qt.to(torch.float32)
        qt = qt.contiguous()
        nbytes = qt.numel() * qt.element_size()
        if qt.is_pinned():
            return qt.to(self.device, non_blocking=True).to(torch.float32)
        if nbytes < (64 << 10):                           # single queries: the driver's inline copy is as fast
            return qt.to(self.device).to(torch.float32)
        with self._stage_lock:                             # retrievers run from thread pools (base.py:82-96)
            st = getattr(self, "_q_stage", None)
            if st is None or st[0].numel() < nbytes:
                st = self._q_stage = (torch.empty((nbytes,), dtype=torch.uint8).pin_memory(), torch.cuda.Event())
            else:
                st[1].synchronize()                        # the previous transfer out of the buffer is done
            view = st[0][:nbytes].view(qt.dtype).view(qt.shape)
            view.copy_(qt)
            dev = view.to(self.device, non_blocking=True)
            st[1].record(torch.cuda.current_stream(self.device))
        return dev.to(torch.float32)

    def prepare_queries(self, q) -> torch.Tensor:
        qt = torch.as_tensor(q)
        if qt.dim() == 1:
            qt = qt[None, :]
        qt = self._to_device_f32(qt)
        if self.x3:
            return ops.normalize_split3(qt, self.normalize)
        if self.l2:
            return ops.l2_augment(qt, self.dtype, is_query=True, normalize=self.normalize)      # [q | 1]
        return ops.normalize_cast(qt, self.dtype, self.normalize)

    def prepare_plain(self, q) -> torch.Tensor:
        """Queries in the row space of ``rows`` (normalised if the store normalises, storage dtype),
        e.g. for MMR, whatever form ``prepare_queries`` needs for searching."""
        qt = torch.as_tensor(q)
        if qt.dim() == 1:
            qt = qt[None, :]
        qt = qt.to(device=self.device, dtype=torch.float32).contiguous()
        return ops.normalize_cast(qt, torch.float32 if self.x3 else self.dtype, self.normalize)

    def search_device(self, q_prepared: torch.Tensor, k: int):
        if self.x3:
            return ops.dense_topk_x3(self.planes, q_prepared, k, n_rows=self.ntotal)
        if self.l2:
            scores, rows = ops.dense_topk(self.aug, q_prepared, k, n_rows=self.ntotal)
            return ops.l2_distances(scores, q_prepared, self.d), rows                            # squared L2, ascending
        return ops.dense_topk(self.rows, q_prepared, k, n_rows=self.ntotal)

    def capture_search(self, queries: torch.Tensor, k: int):
        """CUDA-graph ``search_device`` for a fixed, already prepared query buffer: returns
        ``(replay, scores, rows)``; refill ``queries`` in place, call ``replay()``."""
        torch.cuda.synchronize()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        scope = ops.WorkspaceScope()               # the graph's own scratch buffers (kept alive with it)
        with scope, torch.cuda.stream(side):
            self.search_device(queries, k)         # the scope's workspaces exist now
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                scores, rows = self.search_device(queries, k)
        torch.cuda.current_stream(self.device).wait_stream(side)

        def replay(_g=g, _scope=scope):
            _g.replay()

        return replay, scores, rows

    def score_phase(self, q_prepared: torch.Tensor, k: int, workspace_clean: bool = False) -> None:
        """Phase 1 of ``search_device``: scoring + per-slice selection into the active workspace."""
        ops.dense_topk_phase(ops.N.PHASE_SCORE, self.aug if self.l2 else self.rows, q_prepared, q_prepared.shape[0], k,
                             n_rows=self.ntotal, workspace_clean=workspace_clean)

    def select_phase(self, q_prepared: torch.Tensor, k: int, out=None):
        """Phase 2: merge of the candidate lists a ``score_phase`` call left in the same workspace."""
        nq = q_prepared.shape[0]
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.float32, device=self.device),
                   torch.empty((nq, k), dtype=torch.int64, device=self.device))
        ops.dense_topk_phase(ops.N.PHASE_SELECT, self.aug if self.l2 else self.rows, None, nq, k, n_rows=self.ntotal, out=out)
        if self.l2:
            ops.l2_distances(out[0], q_prepared, self.d)
        return out

    def capture_search_overlapped(self, queries: torch.Tensor, k: int):
        """Throughput form of ``capture_search`` for back-to-back batches: two alternating slots, the
        scoring of call i+1 (its own high-priority stream) overlaps the selection / merge of call i (a
        second stream) - the merge kernel's small CTAs fit beside the persistent scoring CTAs on every
        SM, so its ~35 us disappear from the step.  Returns ``(replay, finish, outs)``: ``replay()``
        enqueues one search of the (fixed) ``queries`` buffer, ``finish()`` makes the current stream
        wait for everything enqueued so far, ``outs[j]`` = ``(scores, rows)`` of the calls with parity
        j.  Every call still does all of its work; only the order in which the GPU runs it changes."""
        if self.x3:
            raise ValueError("the overlapped form is offered for bf16 / fp16 / fp32-FMA stores")
        dev = self.device
        torch.cuda.synchronize(dev)
        s_score = torch.cuda.Stream(dev, priority=-1)
        s_sel = torch.cuda.Stream(dev, priority=0)
        cur = torch.cuda.current_stream(dev)
        slots = []
        for _ in range(2):
            scope = ops.WorkspaceScope()
            out = (torch.empty((queries.shape[0], k), dtype=torch.float32, device=dev),
                   torch.empty((queries.shape[0], k), dtype=torch.int64, device=dev))
            s_score.wait_stream(cur)
            with scope, torch.cuda.stream(s_score):
                self.score_phase(queries, k)               # warm-up: the scope's workspace exists now
                s_score.synchronize()
            s_sel.wait_stream(s_score)
            with scope, torch.cuda.stream(s_sel):
                self.select_phase(queries, k, out)         # ... and leaves thresholds / rungs reset behind its merge
                s_sel.synchronize()
                g_sel = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_sel, stream=s_sel):
                    self.select_phase(queries, k, out)
            with scope, torch.cuda.stream(s_score):
                # the scoring graph is the scoring kernels only: every replay follows a select on the slot
                g_score = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_score, stream=s_score):
                    self.score_phase(queries, k, workspace_clean=True)
            slots.append({"scope": scope, "out": out, "g_score": g_score, "g_sel": g_sel,
                          "scored": torch.cuda.Event(), "selected": torch.cuda.Event()})
        torch.cuda.synchronize(dev)
        state = {"i": 0}

        def replay():
            sl = slots[state["i"] & 1]
            if state["i"] < 2:
                s_score.wait_stream(torch.cuda.current_stream(dev))        # inputs written on the caller's stream
            state["i"] += 1
            s_score.wait_event(sl["selected"])             # the slot's lists have been merged (two calls ago)
            with torch.cuda.stream(s_score):
                sl["g_score"].replay()
                sl["scored"].record(s_score)
            s_sel.wait_event(sl["scored"])
            with torch.cuda.stream(s_sel):
                sl["g_sel"].replay()
                sl["selected"].record(s_sel)
            return sl["out"]

        def finish():
            c = torch.cuda.current_stream(dev)
            c.wait_stream(s_score); c.wait_stream(s_sel)

        return replay, finish, [sl["out"] for sl in slots]

    def search(self, q, k: int):
        """faiss-style: numpy in, ``(D float32 [nq,k], I int64 [nq,k])`` numpy out."""
        D, I = self.search_device(self.prepare_queries(q), k)
        return D.cpu().numpy(), I.cpu().numpy()


class SearchPipeline:
    """Throughput-oriented batched search over HOST buffers: the host->device copy of batch i+1 and
    the device->host copy of batch i-1 overlap the kernels of batch i (three CUDA streams, `depth`
    staging slots, events between them).  Every batch still pays its own H2D and D2H; only their
    latency is hidden.  The kernels of a slot (query normalise + cast, seeding, scoring, merge) are
    replayed from a CUDA graph, which removes the launch gaps between them.

        pipe = store.pipeline(nq=1024, k=100)
        t = pipe.submit(q_host_fp32)          # pinned [nq,d] fp32 tensor (or numpy array)
        scores, rows = pipe.result(t)         # pinned host tensors, valid until the slot is reused
    """

    def __init__(self, index: "FlatIndexB200", nq: int, k: int, depth: int = 2, graph: bool = True):
        self.index, self.nq, self.k, self.depth = index, nq, k, depth
        dev = index.device
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.slots = []
        for _ in range(depth):
            self.slots.append({
                "q32": torch.zeros((nq, index.d), dtype=torch.float32, device=dev),
                "h_scores": torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                "h_rows": torch.empty((nq, k), dtype=torch.int64).pin_memory(),
                "copied_in": torch.cuda.Event(), "computed": torch.cuda.Event(), "copied_out": torch.cuda.Event(),
                "busy": False, "graph": None,
            })
        self.n_submitted = 0
        self.use_graph = graph
        self._captured_for = None
        self._scope = None
        self._eager_scope = ops.WorkspaceScope()
        if graph:
            self._capture()

    def _index_state(self):
        rows = self.index.planes if self.index.x3 else (self.index.aug if self.index.l2 else self.index.rows)
        return (rows.data_ptr(), self.index.ntotal)

    def _capture(self) -> None:
        """One CUDA graph per staging slot: query normalise + cast, seeding, scoring, merge.  The
        graphs hold the index's row pointer and size, so they are re-captured when the index has
        grown or shrunk since (checked on every submit)."""
        dev = self.index.device
        try:
            torch.cuda.synchronize(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            # scratch buffers owned by this pipeline: the slots replay one after the other on s_cmp, so
            # they may share them, but nothing outside the pipeline can touch or replace them
            self._scope = ops.WorkspaceScope()
            with self._scope, torch.cuda.stream(side):
                for slot in self.slots:                    # the scope's workspaces exist now
                    self.index.search_device(self.index.prepare_queries(slot["q32"]), self.k)
                side.synchronize()
                for slot in self.slots:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        slot["scores"], slot["rows"] = self.index.search_device(
                            self.index.prepare_queries(slot["q32"]), self.k)
                    slot["graph"] = g
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self._captured_for = self._index_state()
        except Exception as exc:  # noqa: BLE001 - capture is an optimisation; the eager path is the same work
            import warnings
            warnings.warn(f"SearchPipeline: CUDA graph capture failed ({type(exc).__name__}: {exc}); launching eagerly")
            torch.cuda.synchronize(dev)
            for slot in self.slots:
                slot["graph"] = None
            self.use_graph = False

    def submit(self, q_host) -> int:
        ticket = self.n_submitted
        slot = self.slots[ticket % self.depth]
        if slot["busy"]:
            raise RuntimeError("pipeline slot still in use: call result() for older tickets first")
        qh = torch.as_tensor(q_host)
        if qh.shape != (self.nq, self.index.d) or qh.dtype != torch.float32:
            raise ValueError(f"expected float32 [{self.nq},{self.index.d}] queries")
        if self.use_graph and self._captured_for != self._index_state():
            if any(s["busy"] for s in self.slots):
                raise RuntimeError("the index changed while batches are in flight: collect them first")
            self._capture()
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(slot["computed"])            # previous use of this slot's q32 is done
            slot["q32"].copy_(qh, non_blocking=True)
            slot["copied_in"].record(self.s_in)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(slot["copied_in"])
            self.s_cmp.wait_event(slot["copied_out"])         # previous results of this slot have left
            if slot["graph"] is not None:
                slot["graph"].replay()
            else:
                with self._eager_scope:
                    slot["scores"], slot["rows"] = self.index.search_device(self.index.prepare_queries(slot["q32"]), self.k)
            slot["computed"].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["computed"])
            slot["h_scores"].copy_(slot["scores"], non_blocking=True)
            slot["h_rows"].copy_(slot["rows"], non_blocking=True)
            if slot["graph"] is None:
                slot["scores"].record_stream(self.s_out); slot["rows"].record_stream(self.s_out)
            slot["copied_out"].record(self.s_out)
        slot["busy"] = True
        self.n_submitted += 1
        return ticket

    def result(self, ticket: int):
        slot = self.slots[ticket % self.depth]
        slot["copied_out"].synchronize()
        slot["busy"] = False
        return slot["h_scores"], slot["h_rows"]


class B200VectorStore(VectorStore):
    def __init__(self, embedding, index: Optional[FlatIndexB200] = None, index_type: str = "flat",
                 metric: str = "cosine", normalize_L2: bool = False, dtype="float32", device="cuda",
                 **kwargs: Any):
        super().__init__(**kwargs)
        self.embedding = embedding
        self.index_type = index_type
        self.metric = metric
        self.normalize_L2 = normalize_L2
        self.index = index
        self.dtype = _as_dtype(dtype)
        self.x3 = _is_x3(dtype)
        # dtype="float32" (the default, the reference's precision): fp32 master rows, searched on the
        # tensor cores through three bf16 planes whenever the dimension allows (d % 64 == 0, scores within
        # 2e-6 of the fp32 FMA result at 8x its speed); "float32_exact" pins the CUDA-core fp32 FMA path
        self.auto_x3 = (not self.x3) and (not _is_exact_fp32(dtype)) and self.dtype == torch.float32
        self.dtype_name = ("float32x3" if self.x3 else "float32_exact" if _is_exact_fp32(dtype)
                           else str(self.dtype).replace("torch.", ""))
        self.device = torch.device(device)
        self.docstore: dict[str, Document] = {}
        self.index_to_docstore_id: dict[int, str] = {}
        self._version = 0            # bumped by every add / delete (row -> Document caches key on it)

    # ---- index management --------------------------------------------------------------------
    def _get_dimension(self) -> int:
        if self.index is not None:
            return self.index.d
        return len(self.embedding.embed_query("test"))

    def _create_index(self, dimension: int) -> FlatIndexB200:
        if self.metric not in ("cosine", "ip", "l2"):
            raise ValueError(f"unsupported metric: {self.metric}")
        if self.index_type != "flat":
            raise ValueError(f"unsupported index type: {self.index_type} (B200VectorStore is exact/flat only)")
        if self.metric == "l2" and self.x3:
            raise ValueError("dtype 'float32x3' is offered for the inner-product / cosine metrics only")
        if self.x3 and dimension % 64 != 0:
            raise ValueError("dtype 'float32x3' needs an embedding dimension that is a multiple of 64")
        x3 = self.x3 or (self.auto_x3 and dimension % 64 == 0 and self.metric != "l2")
        return FlatIndexB200(dimension, self.dtype, self.device, self._normalizes(), x3=x3,
                             l2=self.metric == "l2")

    def _normalizes(self) -> bool:
        return bool(self.normalize_L2 or self.metric == "cosine")

    @property
    def ntotal(self) -> int:
        return 0 if self.index is None else self.index.ntotal

    def add_texts(self, texts: List[str], metadatas: Optional[List[dict]] = None, *,
                  ids: Optional[List[str]] = None, **kwargs: Any) -> List[str]:
        texts = list(texts)
        if not texts:
            return []
        vectors = np.asarray(self.embedding.embed_documents(texts), dtype=np.float32)
        return self.add_embeddings(texts, vectors, metadatas, ids=ids)

    def add_embeddings(self, texts: Sequence[str], vectors, metadatas: Optional[List[dict]] = None, *,
                       ids: Optional[List[str]] = None) -> List[str]:
        """Append pre-computed embeddings (fp32 ``[n,d]`` numpy or tensor, host or device)."""
        texts = list(texts)
        if getattr(vectors, "ndim", 2) != 2 or int(vectors.shape[0]) != len(texts):
            raise ValueError(f"number of vectors ({tuple(vectors.shape)}) must match number of texts ({len(texts)})")
        if self.index is None:
            self.index = self._create_index(int(vectors.shape[1]))
        if ids is None:
            ids = [str(uuid.uuid4()) for _ in texts]
        elif len(ids) != len(texts):
            raise ValueError("number of ids must match number of texts")
        if metadatas is None:
            metadatas = [{} for _ in texts]
        elif len(metadatas) != len(texts):
            raise ValueError("number of metadatas must match number of texts")
        start = self.index.ntotal
        self.index.add(vectors)
        for i, (text, meta, doc_id) in enumerate(zip(texts, metadatas, ids)):
            self.docstore[doc_id] = Document(content=text, metadata=meta, id=doc_id)
            self.index_to_docstore_id[start + i] = doc_id
        self._version += 1
        return list(ids)

    # ---- search ------------------------------------------------------------------------------
    def similarity_search(self, query: str, k: int = 4, **kwargs: Any) -> List[Document]:
        return [doc for doc, _ in self.similarity_search_with_score(query, k, **kwargs)]

    def similarity_search_with_score(self, query: str, k: int = 4, **kwargs: Any) -> List[Tuple[Document, float]]:
        if self.ntotal == 0:
            return []
        return self.similarity_search_by_vector_with_score(self.embedding.embed_query(query), k, **kwargs)

    def similarity_search_by_vector(self, embedding: List[float], k: int = 4, **kwargs: Any) -> List[Document]:
        return [doc for doc, _ in self.similarity_search_by_vector_with_score(embedding, k, **kwargs)]

    def similarity_search_by_vector_with_score(self, embedding: List[float], k: int = 4, **kwargs: Any
                                               ) -> List[Tuple[Document, float]]:
        if self.ntotal == 0:
            return []
        k = min(k, self.ntotal)
        D, I = self.index.search(np.asarray([embedding], dtype=np.float32), k)
        out = []
        for score, row in zip(D[0], I[0]):
            if row == -1:
                continue
            out.append((self.docstore[self.index_to_docstore_id[int(row)]], float(score)))
        return out

    def search_batch(self, queries, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Batched search: ``queries`` fp32 ``[nq,d]`` (numpy / host tensor / device tensor) ->
        device tensors ``(scores float32 [nq,k], rows int64 [nq,k])``; -1 rows pad when k > ntotal."""
        if self.index is None:
            raise ValueError("the store is empty")
        return self.index.search_device(self.index.prepare_queries(queries), k)

    def pipeline(self, nq: int, k: int, depth: int = 2) -> SearchPipeline:
        """Double-buffered host-in / host-out batched search (see ``SearchPipeline``)."""
        if self.index is None:
            raise ValueError("the store is empty")
        return SearchPipeline(self.index, nq, min(k, max(self.ntotal, 1)), depth)

    def similarity_search_batch(self, queries: List[str], k: int = 4) -> List[List[Tuple[Document, float]]]:
        if self.ntotal == 0:
            return [[] for _ in queries]
        vecs = np.asarray(self.embedding.embed_documents(list(queries)), dtype=np.float32)
        D, I = self.search_batch(vecs, min(k, self.ntotal))
        D = D.cpu().numpy(); I = I.cpu().numpy()
        return [[(self.docstore[self.index_to_docstore_id[int(r)]], float(s)) for s, r in zip(Dq, Iq) if r != -1]
                for Dq, Iq in zip(D, I)]

    def self_join(self, k: int = 10, min_score: Optional[float] = None, batch: int = 4096):
        """All-pairs nearest neighbours inside the store (the cosine self-join the reference does
        with an O(n^2) host matrix for entity dedup / event KNN: Base_Neo4j.py:538-590 uses
        sklearn cosine >= 0.95, event_graphrag_neo4j.py:637-648 a top-10 KNN with cutoff 0.85).
        Every stored row is used as a query against the whole matrix, ``batch`` rows at a time.
        Returns ``(scores float32 [n,k], rows int64 [n,k])`` on the device with the self match
        removed; entries below ``min_score`` are set to (-inf, -1)."""
        if self.ntotal == 0:
            raise ValueError("the store is empty")
        n = self.ntotal
        kk = min(k + 1, n)
        out_s = torch.full((n, k), float("-inf"), dtype=torch.float32, device=self.device)
        out_r = torch.full((n, k), -1, dtype=torch.int64, device=self.device)
        for s0 in range(0, n, batch):
            e0 = min(n, s0 + batch)
            if self.index.x3:
                sc, rw = ops.dense_topk_x3(self.index.planes, self.index.planes[s0:e0].contiguous(), kk, n_rows=n)
            else:
                sc, rw = ops.dense_topk(self.index.rows, self.index.rows[s0:e0].contiguous(), kk, n_rows=n)
            me = torch.arange(s0, e0, device=self.device)[:, None]
            keep = rw != me
            # drop exactly one entry per row: the self match if present, else the last one
            none = keep.all(dim=1)
            keep[none, kk - 1] = False
            idx = keep.float().argsort(dim=1, descending=True, stable=True)[:, :kk - 1]
            sc = sc.gather(1, idx); rw = rw.gather(1, idx)
            if min_score is not None:
                bad = sc < min_score
                sc = sc.masked_fill(bad, float("-inf")); rw = rw.masked_fill(bad, -1)
            out_s[s0:e0, :kk - 1] = sc
            out_r[s0:e0, :kk - 1] = rw
        return out_s, out_r

    def rows_to_documents(self, rows: Sequence[int]) -> List[Document]:
        return [self.docstore[self.index_to_docstore_id[int(r)]] for r in rows if r != -1]

    # ---- maximal marginal relevance --------------------------------------------------------------
    def max_marginal_relevance_search(self, query: str, k: int = 4, fetch_k: int = 20,
                                      lambda_mult: float = 0.5, **kwargs: Any) -> List[Document]:
        if self.ntotal == 0:
            return []
        return self.max_marginal_relevance_search_by_vector(self.embedding.embed_query(query), k, fetch_k,
                                                            lambda_mult, **kwargs)

    def max_marginal_relevance_search_by_vector(self, embedding: List[float], k: int = 4, fetch_k: int = 20,
                                                lambda_mult: float = 0.5, **kwargs: Any) -> List[Document]:
        if self.ntotal == 0:
            return []
        fetch = min(fetch_k, self.ntotal)
        q = self.index.prepare_queries(np.asarray([embedding], dtype=np.float32))
        _, cand = self.index.search_device(q, fetch)
        if k >= fetch:
            return self.rows_to_documents(cand[0].tolist())
        if self.index.x3 or self.index.l2:      # MMR works on the plain (fp32 master / un-augmented) rows
            q = self.index.prepare_plain(np.asarray([embedding], dtype=np.float32))
        sel = ops.mmr_select(self.index.rows, q, cand.contiguous(), k, lambda_mult, n_rows=self.ntotal)
        cand_h = cand[0].tolist()
        return self.rows_to_documents([cand_h[j] for j in sel[0].tolist() if j >= 0])

    # ---- maintenance ---------------------------------------------------------------------------
    def delete(self, ids: Optional[List[str]] = None, **kwargs: Any) -> Optional[bool]:
        self._version += 1
        if ids is None:
            self.docstore.clear()
            self.index_to_docstore_id.clear()
            if self.index is not None:
                self.index.reset()
            return True
        if not ids:
            return True
        for doc_id in ids:
            if doc_id not in self.docstore:
                return False
        drop = set(ids)
        row_of = {doc_id: row for row, doc_id in self.index_to_docstore_id.items()}
        keep_ids = [doc_id for doc_id in self.docstore if doc_id not in drop]
        keep_rows = torch.tensor([row_of[d] for d in keep_ids], dtype=torch.int64, device=self.device)
        survivors = self.index.rows.index_select(0, keep_rows) if keep_ids else None
        kept_docs = [self.docstore[d] for d in keep_ids]
        self.docstore.clear()
        self.index_to_docstore_id.clear()
        self.index.reset()
        if keep_ids:
            self.index.add_prepared(survivors)
            for row, doc in enumerate(kept_docs):
                self.docstore[doc.id] = doc
                self.index_to_docstore_id[row] = doc.id
        return True

    def get_by_ids(self, ids: Sequence[str], /) -> List[Document]:
        return [self.docstore[i] for i in ids if i in self.docstore]

    def _select_relevance_score_fn(self) -> Callable[[float], float]:
        if self.metric == "cosine" or self.normalize_L2:
            return self._cosine_relevance_score_fn
        if self.metric == "l2":
            return self._euclidean_relevance_score_fn
        if self.metric == "ip":
            return self._max_inner_product_relevance_score_fn
        raise ValueError(f"unsupported metric: {self.metric}")

    # ---- persistence -----------------------------------------------------------------------------
    def save_local(self, folder_path: str, index_name: str = "index") -> None:
        """``<name>.b200.npy`` (the row matrix; bf16 stored as its uint16 bit pattern) +
        ``<name>.pkl`` with the same sidecar keys the reference pickles (:441-450) plus the dtype."""
        os.makedirs(folder_path, exist_ok=True)
        if self.index is not None:
            rows = self.index.rows[:self.index.ntotal]
            host = (rows.view(torch.int16) if self.dtype == torch.bfloat16 else rows).cpu().numpy()
            np.save(os.path.join(folder_path, f"{index_name}.b200.npy"), host)
        side = {"docstore": self.docstore, "index_to_docstore_id": self.index_to_docstore_id,
                "index_type": self.index_type, "metric": self.metric, "normalize_L2": self.normalize_L2,
                "dtype": self.dtype_name}
        with open(os.path.join(folder_path, f"{index_name}.pkl"), "wb") as fh:
            pickle.dump(side, fh)

    @classmethod
    def load_local(cls, folder_path: str, embeddings, index_name: str = "index", **kwargs: Any) -> "B200VectorStore":
        from ....formats import load_reference_sidecar, read_faiss_flat
        side = load_reference_sidecar(os.path.join(folder_path, f"{index_name}.pkl"))
        path = os.path.join(folder_path, f"{index_name}.b200.npy")
        faiss_path = os.path.join(folder_path, f"{index_name}.faiss")
        if os.path.exists(path) and "dtype" in side:
            kwargs["dtype"] = side["dtype"]          # the row file is in exactly this storage type
        else:
            kwargs.setdefault("dtype", "float32")
        store = cls(embedding=embeddings, index_type=side["index_type"], metric=side["metric"],
                    normalize_L2=side["normalize_L2"], **kwargs)
        n_side = len(side["index_to_docstore_id"])
        if os.path.exists(path):
            arr = np.load(path)
            want = {torch.bfloat16: (np.int16, np.uint16), torch.float16: (np.float16,), torch.float32: (np.float32,)}[store.dtype]
            if arr.ndim != 2 or arr.dtype.type not in want:
                raise ValueError(f"{path}: holds {arr.dtype} {arr.shape}, the sidecar says dtype {side.get('dtype')}")
            host = torch.from_numpy(arr)
            if store.dtype == torch.bfloat16:
                host = host.view(torch.bfloat16)
            store.index = store._create_index(int(host.shape[1]))
            store.index.add_prepared(host.to(store.device))
        elif os.path.exists(faiss_path):
            # a folder written by the reference's FaissVectorStore.save_local (:432-450): fp32 rows of
            # a flat index, ALREADY normalised by the reference when its metric was cosine (:178) - they
            # are taken over as stored (normalising again would move them by an ulp for no reason);
            # `dtype=` chooses the B200 storage type, i.e. at most a cast happens here
            rows, _ = read_faiss_flat(faiss_path)
            store.index = store._create_index(int(rows.shape[1]))
            prepared = torch.from_numpy(np.ascontiguousarray(rows)).to(store.device)
            store.index.add_prepared(prepared if store.dtype == torch.float32 else prepared.to(store.dtype))
        # the sidecar must describe exactly the rows that were loaded: a truncated or mismatched folder
        # fails here, not with a KeyError (or a wrong document) at search time
        n_rows = 0 if store.index is None else store.index.ntotal
        if n_rows != n_side:
            raise ValueError(f"{folder_path}: the row file holds {n_rows} vectors but the sidecar maps {n_side} rows")
        if n_side and sorted(side["index_to_docstore_id"]) != list(range(n_side)):
            raise ValueError(f"{folder_path}: index_to_docstore_id is not a dense 0..{n_side - 1} row map")
        missing = [i for i in side["index_to_docstore_id"].values() if i not in side["docstore"]]
        if missing:
            raise ValueError(f"{folder_path}: {len(missing)} mapped ids are missing from the docstore (first: {missing[0]!r})")
        store.docstore = side["docstore"]
        store.index_to_docstore_id = side["index_to_docstore_id"]
        return store

    @classmethod
    def from_texts(cls, texts: List[str], embedding, metadatas: Optional[List[dict]] = None, *,
                   ids: Optional[List[str]] = None, **kwargs: Any) -> "B200VectorStore":
        store = cls(embedding=embedding, **kwargs)
        store.add_texts(texts, metadatas=metadatas, ids=ids)
        return store

    @classmethod
    def from_embeddings(cls, texts: Sequence[str], vectors, embedding=None, metadatas=None, *, ids=None,
                        **kwargs: Any) -> "B200VectorStore":
        store = cls(embedding=embedding, **kwargs)
        store.add_embeddings(texts, vectors, metadatas, ids=ids)
        return store
