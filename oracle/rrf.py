"""Oracle (test infrastructure): reciprocal-rank fusion on integer document keys.

Restates ``RRFusion.fuse`` (/root/reference ``core/utils/Fusion.py:45-76``):

* ``:47-49``  rank is re-assigned ``i + 1`` by position inside each retriever's list
* ``:55-61``  ``rrf[key] += 1.0 / (k + rank)`` walking the lists in retriever order; the key is the
  document *content*; ``document_map[key]`` keeps the last document seen for a key
* ``:64``     ``sorted(items, key=score, reverse=True)`` - stable, so equal scores keep
  first-insertion order
* ``:68-74``  slice ``top_k``, ranks ``1..``

Pinned: ``tests/test_oracle_rrf.py`` runs the reference's own class live (``oracle/ref_loader.py``)
and the committed ``tests/golden/rrf_*.json`` vectors were produced by it
(``oracle/gen_golden.py``).  Here keys are integers (canonical document index); the host layer
maps content strings to such keys before fusing.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

__all__ = ["rrf_fuse_ids", "rrf_fuse_rows"]


def rrf_fuse_ids(lists: Sequence[Sequence[int]], top_k: int, k: float = 60.0
                 ) -> Tuple[List[int], List[float]]:
    """Fuse ranked id lists; negative ids are padding and are skipped *without* consuming a rank
    position only if they trail the list (the GPU layout pads short lists with -1 at the end)."""
    acc = {}
    for ranked in lists:
        for pos, key in enumerate(ranked):
            if key < 0:
                continue
            acc[key] = acc.get(key, 0.0) + 1.0 / (k + (pos + 1))
    order = sorted(acc.items(), key=lambda kv: kv[1], reverse=True)[:top_k]
    return [kv[0] for kv in order], [kv[1] for kv in order]


def rrf_fuse_rows(rows: Sequence[Sequence[int]], row_contents: Sequence[Sequence[str]], top_k: int, k: float = 60.0
                  ) -> Tuple[List[Tuple[int, int]], List[float]]:
    """The hybrid merge as the reference performs it on Documents (``mutipath.py:57-93`` feeding
    ``Fusion.py:45-76``), stated on corpus rows: ``rows[l]`` is retriever l's ranked list of rows into its own
    corpus (negative = padding), ``row_contents[l][row]`` the content string of that row.  Returns, per fused
    entry, ``(list, row)`` of the Document ``document_map`` holds when the walk is over - the last (list,
    position) whose content equals the key (``Fusion.py:61``) - and the fused scores."""
    acc, last = {}, {}
    for l, ranked in enumerate(rows):
        for pos, row in enumerate(ranked):
            if row < 0:
                continue
            content = row_contents[l][row]
            acc[content] = acc.get(content, 0.0) + 1.0 / (k + (pos + 1))
            last[content] = (l, row)
    order = sorted(acc.items(), key=lambda kv: kv[1], reverse=True)[:top_k]
    return [last[c] for c, _ in order], [sc for _, sc in order]
