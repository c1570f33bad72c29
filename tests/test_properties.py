"""Property tests (hypothesis) of the test infrastructure and host logic that everything else leans
on: the RRF oracle against the reference's own RRFusion executed live, the tie-aware comparator,
shard bounds, and the FAISS flat-file reader/writer."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import dense as odense
from oracle import ref_loader
from oracle import rrf as orrf
from oracle.compare import check_topk_against_scores

SET = settings(max_examples=60, deadline=None, derandomize=True, database=None,
               suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


@SET
@given(st.lists(st.lists(st.integers(0, 30), min_size=0, max_size=12, unique=True), min_size=1, max_size=4),
       st.integers(1, 15))
def test_rrf_oracle_equals_live_reference_rrfusion(lists, top_k):
    """Integer-key oracle vs the reference's RRFusion.fuse over Documents whose content is the key
    (Fusion.py:45-76): same keys, same order (stable first-insertion ties), bit-equal fp64 scores."""
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    ns = ref_loader.load()
    results = [[ns.RetrievalResult(document=ns.Document(content=f"c{key}", metadata={}, id=f"{li}-{key}"),
                                   score=0.0, rank=pos + 1) for pos, key in enumerate(ranked)]
               for li, ranked in enumerate(lists)]
    fused = ns.RRFusion().fuse(results, top_k)
    ids, scores = orrf.rrf_fuse_ids(lists, top_k)
    assert [r.document.content for r in fused] == [f"c{k}" for k in ids]
    assert [np.float64(r.score).tobytes() for r in fused] == [np.float64(s).tobytes() for s in scores]
    assert [r.rank for r in fused] == list(range(1, len(ids) + 1))


@SET
@given(st.integers(1, 400), st.integers(1, 9))
def test_shard_bounds_tile_the_row_range(n, world):
    from rag_arc_b200.sharded import shard_bounds
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a <= b for a, b in spans) and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    per = -(-n // world)
    assert all(b - a <= per for a, b in spans)


@SET
@given(st.integers(2, 60), st.integers(1, 10), st.integers(0, 2 ** 31 - 1))
def test_comparator_accepts_tie_permutations_and_rejects_wrong_rows(n, k, seed):
    rng = np.random.default_rng(seed)
    scores = np.round(rng.standard_normal(n), 1)              # many exact ties
    k = min(k, n)
    order = np.lexsort((np.arange(n), -scores))[:k]
    check_topk_against_scores(order, scores[order], scores, k)
    # any other choice inside the boundary tie group is as good
    kth = scores[order[-1]]
    tied_out = [i for i in np.flatnonzero(scores == kth) if i not in set(order.tolist())]
    if tied_out:
        alt = order.copy(); alt[-1] = tied_out[0]
        check_topk_against_scores(alt, scores[alt], scores, k)
    # a row that scores strictly below something left out must be rejected
    worse = [i for i in range(n) if scores[i] < kth]
    if worse:
        bad = order.copy(); bad[-1] = worse[0]
        with pytest.raises(AssertionError):
            check_topk_against_scores(bad, scores[bad], scores, k)


@SET
@given(st.integers(0, 40), st.integers(1, 17), st.sampled_from(["ip", "l2"]), st.integers(0, 2 ** 31 - 1))
def test_faiss_flat_files_round_trip(tmp_path_factory, n, d, metric, seed):
    from rag_arc_b200 import formats
    rows = np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32)
    path = str(tmp_path_factory.mktemp("ff") / "i.faiss")
    formats.write_faiss_flat(path, rows, metric)
    got, m = formats.read_faiss_flat(path)
    assert m == metric and got.shape == (n, d) and np.array_equal(got, rows)
    assert os.path.getsize(path) == 4 + 4 + 8 + 16 + 1 + 4 + 8 + 4 * n * d


@SET
@given(st.integers(1, 50), st.integers(1, 12), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
def test_dense_oracle_blocked_search_equals_one_shot(n, d, k, seed):
    """The oracle's blocked scan + merge (how it handles 1M rows in bounded memory) against a
    single-block scan, including duplicate rows and k > n padding.  BLAS may round the same dot
    product differently for different block shapes, so the two are compared the tie-aware way:
    both must be valid top-k lists of the fp64 scores and agree on the scores to 1e-6."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(np.float32)
    if n > 3:
        X[n - 1] = X[0]
    Q = rng.standard_normal((3, d)).astype(np.float32)
    D1, I1 = odense.flat_ip_search(X, Q, k, block=7)
    D2, I2 = odense.flat_ip_search(X, Q, k, block=1 << 20)
    S = Q.astype(np.float64) @ X.astype(np.float64).T
    kk = min(k, n)
    for qi in range(3):
        for D, I in ((D1, I1), (D2, I2)):
            check_topk_against_scores(I[qi, :kk], D[qi, :kk], S[qi], kk, rtol=1e-5, atol=1e-6)
    assert np.allclose(D1[:, :kk], D2[:, :kk], rtol=1e-6, atol=1e-6)
    if k > n:
        assert (I1[:, n:] == -1).all() and (I2[:, n:] == -1).all()
