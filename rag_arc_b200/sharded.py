"""Row-sharded exact dense search over the GPUs of one node (one process per GPU).

The flat index shards naturally: GPU g holds the contiguous rows ``[g*ceil(N/G), (g+1)*ceil(N/G))``
of the corpus, every rank scores the full (replicated) query batch against its shard with
``ragarc_dense_topk_keys`` - which already emits packed sortable keys carrying GLOBAL row ids,
sorted per query - and the ``G`` per-shard lists of each query are merged on every rank.  Because
keys order by (score, lowest global row id), the result is bit-identical for any G, including G=1.

Two exchange paths for the ``[nq,k]`` key blocks (``nq*k*8`` bytes per rank: latency-, not
bandwidth-bound on NVSwitch):

* **peer memory** (default when it can be set up): every rank writes its keys into a buffer
  allocated as symmetric memory, one device-side barrier orders the writes, and the merge kernel
  (``ragarc_merge_topk_keys_p2p``) loads the other ranks' lists straight over NVLink through a table
  of peer pointers - the gather is fused into the merge, no collective call on the data path.
  Buffers are double-buffered so that one barrier per search suffices (a rank that passes the
  barrier of search i+1 knows every rank has finished reading search i's buffer).
* **NCCL**: one ``all_gather_into_tensor`` followed by ``ragarc_merge_topk_keys``.

The reference has no distributed code at all; this is the multi-GPU form of
``faiss.IndexFlatIP.search`` (VectorStore_Faiss.py:263).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


class _PeerExchange:
    """Symmetric-memory key buffers + peer pointer tables for one (nq, k) shape."""

    def __init__(self, nq: int, k: int, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.nq, self.k = nq, k
        self.buf = symm_mem.empty((2, nq, k), dtype=torch.int64, device=device)
        grp = group if group is not None else dist.group.WORLD
        self.hdl = symm_mem.rendezvous(self.buf, grp)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        slot_bytes = nq * k * 8
        self.tables = [torch.tensor([p + s * slot_bytes for p in ptrs], dtype=torch.int64, device=device)
                       for s in range(2)]
        self.step = 0

    def exchange_and_merge(self, write_keys, k_out: int):
        slot = self.step & 1
        self.step += 1
        write_keys(self.buf[slot])                 # this rank's keys -> its symmetric buffer
        self.hdl.barrier(channel=slot)             # everybody's keys of this search are in place
        return ops.merge_topk_keys_p2p(self.tables[slot], self.nq, self.k, k_out)


class _OwnerExchange:
    """Query-owner exchange for one (nq, k) shape: rank g owns queries ``[g*nq_per, (g+1)*nq_per)``;
    every rank's merge kernel PUSHES the sorted key row of each query into its owner's inbox
    (symmetric memory, stores over NVLink) and bumps the owner's arrival counter of that query; the
    owner's merge kernel waits per query for ``world`` arrivals and merges the ``world`` lists of its
    own queries out of local memory.  Compared with every rank merging all ``nq`` queries from peer
    memory this divides the cross-shard merge work by ``world``, turns remote loads (a round trip each)
    into fire-and-forget stores, and needs neither a collective nor a barrier kernel between the two
    launches.  (Measured at 2 GPUs the counter variant is no faster than one symmetric-memory barrier
    between the two launches, so the barrier-ordered variant is the default; ``RAGARC_OWNER_SIGNAL=1``
    selects the counters.)
    Two inboxes alternate: a rank can be at most one search ahead of the slowest rank, because its
    own merge of search i+1 needs that rank's rows of search i+1, which are only pushed after that
    rank has finished merging search i."""

    def __init__(self, nq: int, k: int, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(grp), dist.get_rank(grp)
        self.nq, self.k = nq, k
        self.nq_per = (nq + self.world - 1) // self.world
        self.q_lo = min(nq, self.rank * self.nq_per)
        self.q_hi = min(nq, self.q_lo + self.nq_per)
        self.signalled = os.environ.get("RAGARC_OWNER_SIGNAL", "0") == "1"
        self.words = ops.inbox_words(self.world, self.nq_per, k)
        self.inbox = symm_mem.empty((2, self.words), dtype=torch.int64, device=device)
        self.inbox.zero_()
        self.hdl = symm_mem.rendezvous(self.inbox, grp)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.tables = [torch.tensor([p + s * self.words * 8 for p in ptrs], dtype=torch.int64, device=device)
                       for s in range(2)]
        self.status = torch.zeros((1,), dtype=torch.int32, device=device)
        self.step = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=grp)                      # every inbox is zeroed before anybody pushes

    def timed_out(self) -> bool:
        """True if a merge kernel gave up waiting for a peer's rows since the last call (synchronises)."""
        bad = bool(int(self.status.item()) != 0)
        self.status.zero_()
        return bad

    def push_and_merge(self, push, k_out: int):
        slot = self.step & 1
        self.step += 1
        push(self.tables[slot], self.signalled)      # this rank's key rows -> the owners' inboxes
        return self.merge_slot(slot, k_out)

    def merge_slot(self, slot: int, k_out: int):
        """Owner half for inbox ``slot``: wait until every rank's rows of this search have landed (arrival
        counters inside the merge kernel, or one symmetric-memory barrier), merge the owned queries."""
        n_own = self.q_hi - self.q_lo
        if self.signalled:
            return ops.merge_topk_inbox(self.inbox[slot], self.world, self.nq_per, n_own, self.k, k_out, self.status)
        self.hdl.barrier(channel=slot)               # all rows of this search have landed
        keys = self.inbox[slot][:self.world * self.nq_per * self.k].view(self.world, self.nq_per, self.k)
        scores, rows = ops.merge_topk_keys(keys, k_out)
        return scores[:n_own], rows[:n_own]


class ShardedFlatIndex:
    def __init__(self, rows: torch.Tensor, id_base: int, n_rows: Optional[int] = None, group=None,
                 exchange: str = "auto"):
        """rows: this rank's shard ``[n_local(+spare), d]`` (normalised, storage dtype, on this
        rank's GPU); id_base: global row id of ``rows[0]``; exchange: "auto" | "peer" | "nccl"."""
        self.rows = rows
        self.id_base = int(id_base)
        self.n_local = rows.shape[0] if n_rows is None else int(n_rows)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.exchange = os.environ.get("RAGARC_EXCHANGE", exchange)
        self._gather_buf = None
        self._peer = {}
        self._owner = {}
        self._peer_failed = False
        self.exchange_used = "none" if self.world == 1 else "nccl"

    def _peer_exchange(self, nq: int, k: int):
        key = (nq, k)
        px = self._peer.get(key)
        if px is None and not self._peer_failed:
            err = None
            try:
                px = _PeerExchange(nq, k, self.rows.device, self.group)
            except Exception as exc:  # noqa: BLE001 - symmetric memory not available on this system
                if self.exchange == "peer":
                    raise
                err, px = exc, None
            # all ranks must agree on the path: if any rank failed, everyone uses NCCL
            ok = torch.tensor([1 if px is not None else 0], device=self.rows.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                px, self._peer_failed = None, True
                if self.rank == 0:
                    print(f"[rag_arc_b200.sharded] peer-memory exchange unavailable "
                          f"({type(err).__name__ if err else 'peer rank'}: {err}); using NCCL all-gather")
            self._peer[key] = px
        return px

    def owned_range(self, nq: int) -> Tuple[int, int]:
        """Queries ``[lo, hi)`` of an ``nq``-query batch whose results ``search_owned`` leaves on this rank."""
        per = (nq + self.world - 1) // self.world
        lo = min(nq, self.rank * per)
        return lo, min(nq, lo + per)

    def search_owned(self, queries: torch.Tensor, k: int):
        """Like ``search``, but the merged result of every query ends up only on the rank that owns
        it (``owned_range``): returns ``(scores float32 [n_own,k], global rows int64 [n_own,k])``.
        Every rank still receives and scores the full batch against its shard; what is partitioned is
        the cross-shard merge and the result (a host gathers the slices, or each rank answers for
        its queries).  Needs peer memory; falls back to ``search`` + slicing otherwise."""
        nq = queries.shape[0]
        lo, hi = self.owned_range(nq)
        if self.world == 1 or self.exchange == "nccl" or self._peer_failed:
            s, r = self.search(queries, k)
            return s[lo:hi], r[lo:hi]
        ox = self._owner.get((nq, k))
        if ox is None:
            err = None
            try:
                ox = _OwnerExchange(nq, k, self.rows.device, self.group)
            except Exception as exc:  # noqa: BLE001 - symmetric memory not available on this system
                if self.exchange == "peer":
                    raise
                err, ox = exc, None
            ok = torch.tensor([1 if ox is not None else 0], device=self.rows.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self._peer_failed = True
                if self.rank == 0:
                    print(f"[rag_arc_b200.sharded] owner exchange unavailable "
                          f"({type(err).__name__ if err else 'peer rank'}: {err}); using NCCL all-gather")
                s, r = self.search(queries, k)
                return s[lo:hi], r[lo:hi]
            self._owner[(nq, k)] = ox
        self.exchange_used = "owner-push" if ox.signalled else "owner-push+barrier"
        return ox.push_and_merge(
            lambda table, signal: ops.dense_topk_keys_push(self.rows, queries, k, self.id_base, table, self.rank,
                                                           ox.nq_per, signal=signal, n_rows=self.n_local), k)

    def capture_owned_overlapped(self, queries: torch.Tensor, k: int):
        """Throughput form of ``capture(owned=True)``: per exchange slot one graph with the scoring phase
        (high-priority stream) and one with everything after it - merge of the local candidate lists
        pushed into the owners' inboxes, the cross-rank ordering, the owner's merge - on a second
        stream, so that the exchange of search i runs under the scoring of search i+1.  Returns
        ``(replay, finish, outs)`` like ``FlatIndexB200.capture_search_overlapped``; every rank must
        call ``replay`` the same number of times."""
        nq = queries.shape[0]
        dev = queries.device
        self.search_owned(queries, k); self.search_owned(queries, k)      # rendezvous, both inbox slots used once
        ox = self._owner.get((nq, k))
        if ox is None:
            raise RuntimeError("the query-owner exchange is not available (no peer memory): use capture(owned=True)")
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)
        s_score = torch.cuda.Stream(dev, priority=-1)
        s_sel = torch.cuda.Stream(dev, priority=0)
        cur = torch.cuda.current_stream(dev)
        N = ops.N
        slots = []
        for slot in range(2):
            scope = ops.WorkspaceScope()

            def score(clean=False):
                ops.dense_topk_phase(N.PHASE_SCORE, self.rows, queries, nq, k, n_rows=self.n_local, workspace_clean=clean)

            def select(slot=slot):
                ops.dense_topk_phase(N.PHASE_SELECT, self.rows, None, nq, k, n_rows=self.n_local, id_base=self.id_base,
                                     inbox_table=ox.tables[slot], rank=self.rank, nq_per_rank=ox.nq_per, signal=ox.signalled)
                return ox.merge_slot(slot, k)

            s_score.wait_stream(cur)
            with scope, torch.cuda.stream(s_score):
                score()
                s_score.synchronize()
            s_sel.wait_stream(s_score)
            with scope, torch.cuda.stream(s_sel):
                select()                                   # eager once: every rank pushes + merges this slot
                s_sel.synchronize()
            dist.barrier(group=self.group)
            with scope, torch.cuda.stream(s_score):
                g_score = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_score, stream=s_score):
                    score(clean=True)                      # the eager select above left the thresholds reset
            with scope, torch.cuda.stream(s_sel):
                g_sel = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_sel, stream=s_sel):
                    out = select()
            slots.append({"scope": scope, "out": out, "g_score": g_score, "g_sel": g_sel,
                          "scored": torch.cuda.Event(), "selected": torch.cuda.Event()})
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)
        state = {"i": 0}

        def replay():
            sl = slots[state["i"] & 1]
            if state["i"] < 2:
                s_score.wait_stream(torch.cuda.current_stream(dev))
            state["i"] += 1
            s_score.wait_event(sl["selected"])
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

        self.exchange_used = ("owner-push" if ox.signalled else "owner-push+barrier") + ", exchange overlapped with the next scoring"
        return replay, finish, [sl["out"] for sl in slots]

    def capture(self, queries: torch.Tensor, k: int, owned: bool = False):
        """CUDA-graph the whole search (scoring, key exchange, merge) for a fixed query buffer:
        returns ``(replay, scores, rows)``; refill ``queries`` in place and call ``replay()``.
        At small shards the step is a handful of ~0.1 ms kernels and host launch overhead shows."""
        search = self.search_owned if owned else self.search
        search(queries, k)                         # warm-up: allocations, peer rendezvous
        search(queries, k)                         # both buffer slots have been used once
        torch.cuda.synchronize()
        side = torch.cuda.Stream(queries.device)
        side.wait_stream(torch.cuda.current_stream(queries.device))
        graphs, outs = [], []
        scope = ops.WorkspaceScope()               # scratch buffers owned by these graphs
        with scope, torch.cuda.stream(side):
            search(queries, k)
            search(queries, k)                     # the scope's workspaces exist now
            side.synchronize()
            for _ in range(2):                     # one graph per exchange-buffer slot
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    outs.append(search(queries, k))
                graphs.append(g)
        torch.cuda.current_stream(queries.device).wait_stream(side)
        state = {"i": 0, "scope": scope}
        scores, rows = outs[0]

        def replay():
            j = state["i"] & 1
            state["i"] += 1
            graphs[j].replay()
            if j == 1:                             # keep one result location for the caller
                scores.copy_(outs[1][0]); rows.copy_(outs[1][1])
            return scores, rows

        return replay, scores, rows

    def search(self, queries: torch.Tensor, k: int):
        """queries: ``[nq,d]`` prepared (normalised, storage dtype), identical on every rank.
        Returns ``(scores float32 [nq,k], global rows int64 [nq,k])`` on every rank."""
        nq = queries.shape[0]
        if self.world == 1:
            keys = ops.dense_topk_keys(self.rows, queries, k, id_base=self.id_base, n_rows=self.n_local)
            return ops.merge_topk_keys(keys.view(1, nq, k), k)
        if self.exchange in ("auto", "peer") and queries.is_cuda:
            px = self._peer_exchange(nq, k)
            if px is not None:
                self.exchange_used = "peer"
                return px.exchange_and_merge(
                    lambda out: ops.dense_topk_keys(self.rows, queries, k, id_base=self.id_base,
                                                    n_rows=self.n_local, out=out), k)
        keys = ops.dense_topk_keys(self.rows, queries, k, id_base=self.id_base, n_rows=self.n_local)
        buf = self._gather_buf
        if buf is None or buf.shape != (self.world, nq, k) or buf.device != keys.device:
            buf = torch.empty((self.world, nq, k), dtype=torch.int64, device=keys.device)
            self._gather_buf = buf
        dist.all_gather_into_tensor(buf.view(-1), keys.view(-1), group=self.group)
        self.exchange_used = "nccl"
        return ops.merge_topk_keys(buf, k)


class ShardedBm25Index:
    """Doc-range sharded BM25: rank g scores the documents
    ``shard_bounds(N, G, g)`` with the GLOBAL idf / average length, the per-shard top-k (fp64 scores,
    global doc ids) are all-gathered (one collective: scores and ids packed into one int64 block of
    ``2*nq*k*8`` bytes per rank) and merged on every rank by ``ragarc_bm25_merge_topk``.  The merged
    result is bit-identical to the single-index one: the score of a document never depends on which
    shard holds it, and ties break towards the lowest global doc id in both.

    The reference scores BM25 on one host only (core/retrieval/bm25.py:297-311)."""

    def __init__(self, full_index, device, group=None):
        """full_index: a ``Bm25Index`` of the WHOLE corpus built on the host (``device=None``) - every
        rank builds or loads the same one, as it needs the global statistics anyway."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_docs = full_index.n_docs
        lo, hi = shard_bounds(full_index.n_docs, self.world, self.rank)
        self.local = full_index.shard(lo, hi, device)
        self.full = full_index
        self._buf = None

    def encode_queries(self, queries):
        return self.local.encode_queries(queries)

    def encode_query_ids(self, qids):
        return self.local.encode_query_ids(qids)

    def search(self, q_terms: torch.Tensor, q_len: torch.Tensor, k: int):
        """Returns ``(scores float64 [nq,k], global doc ids int64 [nq,k])`` on every rank."""
        nq = q_terms.shape[0]
        kk = min(k, self.n_docs)
        k_loc = max(1, min(kk, self.local.n_docs)) if self.local.n_docs > 0 else 1
        if self.local.n_docs > 0:
            s, i = ops.bm25_topk(self.local, q_terms, q_len, k_loc)
        else:                                          # more ranks than documents: nothing to offer
            s = torch.full((nq, k_loc), float("-inf"), dtype=torch.float64, device=q_terms.device)
            i = torch.full((nq, k_loc), -1, dtype=torch.int64, device=q_terms.device)
        # every rank contributes a [2, nq, kk] block (padded when its shard is smaller than kk)
        mine = torch.full((2, nq, kk), -1, dtype=torch.int64, device=q_terms.device)
        mine[0, :, :k_loc] = s.view(torch.int64)
        mine[1, :, :k_loc] = i
        if self.world == 1:
            allb = mine.view(1, 2, nq, kk)
        else:
            if self._buf is None or self._buf.shape != (self.world, 2, nq, kk) or self._buf.device != mine.device:
                self._buf = torch.empty((self.world, 2, nq, kk), dtype=torch.int64, device=mine.device)
            dist.all_gather_into_tensor(self._buf.view(-1), mine.view(-1), group=self.group)
            allb = self._buf
        scores = allb[:, 0].contiguous().view(torch.float64)
        ids = allb[:, 1].contiguous()
        out_s, out_i = ops.bm25_merge_topk(scores, ids, kk)
        if kk < k:                                     # k > number of documents: pad like the single index
            pad_s = torch.full((nq, k - kk), float("-inf"), dtype=torch.float64, device=out_s.device)
            pad_i = torch.full((nq, k - kk), -1, dtype=torch.int64, device=out_i.device)
            out_s, out_i = torch.cat([out_s, pad_s], 1), torch.cat([out_i, pad_i], 1)
        return out_s, out_i


class ShardedSearchPipeline:
    """Host-in / host-out batched search over a ``ShardedFlatIndex`` for throughput: the whole
    per-rank step (query normalise + cast, seeding, scoring, key exchange, merge) of each of two
    staging slots is a CUDA graph, and the H2D copy of batch i+1 and the D2H copy of batch i-1 run
    on their own streams while the graph of batch i executes - every batch still moves its own
    queries in and its own results out.  Every rank must submit the same batches in the same order.

        pipe = ShardedSearchPipeline(sharded, flat_index.prepare_queries, nq, d, k)
        t = pipe.submit(q_host_fp32_pinned)
        scores, rows = pipe.result(t)           # pinned host tensors (valid until the slot is reused)

    The two graphs are bound to the two exchange-buffer slots of the peer-memory path, so while a
    pipeline is in use the index must not be searched through any other route (a full
    ``dist.barrier()`` + device synchronise separates it from earlier work)."""

    def __init__(self, index: ShardedFlatIndex, prepare, nq: int, d: int, k: int, owned: bool = False):
        """owned=True: every rank returns only the results of the queries it owns
        (``index.owned_range(nq)``; query-owner exchange) instead of all ``nq``."""
        self.index, self.nq, self.k = index, nq, k
        self.owned = owned
        search = index.search_owned if owned else index.search
        lo, hi = index.owned_range(nq) if owned else (0, nq)
        self.q_lo, self.q_hi = lo, hi
        n_out = hi - lo
        dev = index.rows.device
        self.dev = dev
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        q32s = [torch.zeros((nq, d), dtype=torch.float32, device=dev) for _ in range(2)]
        for q32 in q32s:                                   # allocations, peer rendezvous
            search(prepare(q32), k)
        torch.cuda.synchronize(dev)
        if index.world > 1:
            dist.barrier(group=index.group)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        self.slots = []
        self._scope = ops.WorkspaceScope()                 # scratch buffers owned by this pipeline's graphs
        with self._scope, torch.cuda.stream(side):
            for q32 in q32s:                               # the scope's workspaces exist now
                search(prepare(q32), k)
            side.synchronize()
            for q32 in q32s:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    scores, rows = search(prepare(q32), k)
                self.slots.append({
                    "q32": q32, "graph": g, "scores": scores, "rows": rows,
                    "h_scores": torch.empty((n_out, k), dtype=torch.float32).pin_memory(),
                    "h_rows": torch.empty((n_out, k), dtype=torch.int64).pin_memory(),
                    "copied_in": torch.cuda.Event(), "computed": torch.cuda.Event(),
                    "copied_out": torch.cuda.Event(), "busy": False,
                })
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if index.world > 1:
            dist.barrier(group=index.group)
        self.n_submitted = 0

    def submit(self, q_host: torch.Tensor) -> int:
        ticket = self.n_submitted
        slot = self.slots[ticket & 1]
        if slot["busy"]:
            raise RuntimeError("pipeline slot still in use: call result() for older tickets first")
        if q_host.shape != slot["q32"].shape or q_host.dtype != torch.float32:
            raise ValueError(f"expected float32 {tuple(slot['q32'].shape)} queries")
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(slot["computed"])        # the slot's previous batch has consumed q32
            slot["q32"].copy_(q_host, non_blocking=True)
            slot["copied_in"].record(self.s_in)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(slot["copied_in"])
            self.s_cmp.wait_event(slot["copied_out"])     # the slot's previous results have left
            slot["graph"].replay()
            slot["computed"].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["computed"])
            slot["h_scores"].copy_(slot["scores"], non_blocking=True)
            slot["h_rows"].copy_(slot["rows"], non_blocking=True)
            slot["copied_out"].record(self.s_out)
        slot["busy"] = True
        self.n_submitted += 1
        return ticket

    def result(self, ticket: int):
        slot = self.slots[ticket & 1]
        slot["copied_out"].synchronize()
        slot["busy"] = False
        return slot["h_scores"], slot["h_rows"]
