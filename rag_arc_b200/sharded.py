"""Row-sharded exact dense search over the GPUs of one node (one process per GPU).

The flat index shards naturally: GPU g holds the contiguous rows ``[g*ceil(N/G), (g+1)*ceil(N/G))``
of the corpus, every rank scores the full (replicated) query batch against its shard with
``ragarc_dense_topk_keys`` - which already emits packed sortable keys carrying GLOBAL row ids -
the ``[nq,k]`` key blocks are exchanged with ONE ``all_gather`` (NCCL over NVLink/NVSwitch;
``nq*k*8`` bytes per rank, latency- not bandwidth-bound) and every rank merges the ``G*k``
candidates per query with ``ragarc_merge_topk_keys``.  Because keys order by (score, lowest global
row id), the result is bit-identical for any G, including G=1.

The reference has no distributed code at all; this is the multi-GPU form of
``faiss.IndexFlatIP.search`` (VectorStore_Faiss.py:263).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


class ShardedFlatIndex:
    def __init__(self, rows: torch.Tensor, id_base: int, n_rows: Optional[int] = None, group=None):
        """rows: this rank's shard ``[n_local(+spare), d]`` (normalised, storage dtype, on this
        rank's GPU); id_base: global row id of ``rows[0]``."""
        self.rows = rows
        self.id_base = int(id_base)
        self.n_local = rows.shape[0] if n_rows is None else int(n_rows)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._gather_buf = None

    def search(self, queries: torch.Tensor, k: int):
        """queries: ``[nq,d]`` prepared (normalised, storage dtype), identical on every rank.
        Returns ``(scores float32 [nq,k], global rows int64 [nq,k])`` on every rank."""
        nq = queries.shape[0]
        keys = ops.dense_topk_keys(self.rows, queries, k, id_base=self.id_base, n_rows=self.n_local)
        if self.world == 1:
            return ops.merge_topk_keys(keys.view(1, nq, k), k)
        buf = self._gather_buf
        if buf is None or buf.shape != (self.world, nq, k) or buf.device != keys.device:
            buf = torch.empty((self.world, nq, k), dtype=torch.int64, device=keys.device)
            self._gather_buf = buf
        dist.all_gather_into_tensor(buf.view(-1), keys.view(-1), group=self.group)
        return ops.merge_topk_keys(buf, k)
