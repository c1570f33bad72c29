"""Tie-aware comparators (test infrastructure).

The reference leaves tie order unspecified (FAISS heap order; numpy's unstable argsort at
core/retrieval/bm25.py:309), so id lists are compared per *score group*: inside a run of
(numerically) equal oracle scores any permutation is accepted, and at the k boundary a group may
be truncated differently.  Everything else must match exactly.
"""
from __future__ import annotations

import numpy as np

__all__ = ["check_topk", "check_topk_against_scores"]


def check_topk(got_ids, got_scores, ref_ids, ref_scores, *, rtol=0.0, atol=0.0, what="topk"):
    """Compare one query's result with the oracle's.

    got/ref_scores are descending.  Scores must agree elementwise within
    ``atol + rtol*|ref|``; ids must agree except inside groups whose oracle scores are closer than
    that tolerance (incl. the boundary group, where members may be swapped for outside ties -
    use ``check_topk_against_scores`` when the full score vector is available).
    """
    got_ids = np.asarray(got_ids)
    ref_ids = np.asarray(ref_ids)
    got_scores = np.asarray(got_scores, dtype=np.float64)
    ref_scores = np.asarray(ref_scores, dtype=np.float64)
    assert got_ids.shape == ref_ids.shape, f"{what}: shape {got_ids.shape} vs {ref_ids.shape}"
    tol = atol + rtol * np.abs(ref_scores)
    bad = np.abs(got_scores - ref_scores) > tol
    assert not bad.any(), (f"{what}: score mismatch at {np.nonzero(bad)[0][:5]}: "
                           f"{got_scores[bad][:5]} vs {ref_scores[bad][:5]}")
    k = len(ref_ids)
    i = 0
    while i < k:
        j = i + 1
        while j < k and abs(ref_scores[j] - ref_scores[j - 1]) <= tol[j]:
            j += 1
        g, r = set(got_ids[i:j].tolist()), set(ref_ids[i:j].tolist())
        if g != r:
            # only legal if the group touches the k boundary (outside ties may have been chosen)
            assert j == k, f"{what}: ids differ in score group [{i},{j}): {sorted(g ^ r)[:8]}"
        i = j
    return True


def check_topk_against_scores(got_ids, got_scores, all_scores, k, *, rtol=0.0, atol=0.0, what="topk"):
    """Strong check when the oracle's full score vector is known: every returned id carries its
    oracle score (within tol), scores are non-increasing, ids are distinct, and nothing outside
    the result beats the k-th returned score by more than tol."""
    got_ids = np.asarray(got_ids)
    got_scores = np.asarray(got_scores, dtype=np.float64)
    all_scores = np.asarray(all_scores, dtype=np.float64)
    n = all_scores.shape[0]
    kk = min(k, n)
    assert len(got_ids) >= kk
    ids = got_ids[:kk]
    assert len(set(ids.tolist())) == kk, f"{what}: duplicate ids"
    assert (ids >= 0).all() and (ids < n).all(), f"{what}: id out of range"
    ref = all_scores[ids]
    tol = atol + rtol * np.abs(ref)
    assert (np.abs(got_scores[:kk] - ref) <= tol).all(), f"{what}: returned score != oracle score"
    assert (np.diff(got_scores[:kk]) <= 0).all(), f"{what}: scores not descending"
    if kk < n:
        rest = np.ones(n, bool)
        rest[ids] = False
        best_outside = all_scores[rest].max()
        kth = ref.min()
        slack = atol + rtol * abs(kth)
        assert best_outside <= kth + 2 * slack, (
            f"{what}: missed a better row: outside {best_outside} > kth {kth}")
    return True
