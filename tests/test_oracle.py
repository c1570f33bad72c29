"""CPU: pin the oracle.  Hand-computed known-answer tests, the golden vectors the reference's own
code produced (oracle/gen_golden.py), and - when /root/reference is present - the live reference."""
import json
import math
import os
from fractions import Fraction

import numpy as np
import pytest

from oracle import bm25 as obm25
from oracle import dense as odense
from oracle import pool as opool
from oracle import ref_loader
from oracle import rrf as orrf
from oracle.compare import check_topk, check_topk_against_scores

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- dense ------------------------------------------------------------------------------------------
def test_normalize_l2_kat_and_zero_rows():
    x = np.array([[3.0, 4.0], [0.0, 0.0], [1e-20, 0.0]], np.float32)
    odense.normalize_L2(x)
    assert np.allclose(x[0], [0.6, 0.8])
    assert (x[1] == 0).all()                      # zero rows untouched (no epsilon)
    assert np.isfinite(x).all()


def test_flat_ip_kat_descending_minus_one_padding_and_ties():
    X = np.array([[1, 0], [0, 1], [1, 1], [1, 0], [-1, 0]], np.float32)
    Q = np.array([[1, 0], [0.5, 0.5]], np.float32)
    D, I = odense.flat_ip_search(X, Q, 7)
    assert I[0].tolist() == [0, 2, 3, 1, 4, -1, -1]            # ties (rows 0,2,3 score 1) by ascending id
    assert D[0, :5].tolist() == [1, 1, 1, 0, -1] and np.isinf(D[0, 5:]).all()
    assert I[1, :3].tolist() == [2, 0, 1]
    idx = odense.IndexFlatIP(2); idx.add(X)
    assert idx.ntotal == 5 and idx.d == 2 and idx.is_trained
    D2, I2 = idx.search(Q, 3)
    assert (I2 == I[:, :3]).all() and (D2 == D[:, :3]).all()
    idx.reset(); assert idx.ntotal == 0


def test_flat_ip_blocked_equals_unblocked_and_f64_adjudicator():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((5000, 32)).astype(np.float32)
    Q = rng.standard_normal((9, 32)).astype(np.float32)
    D1, I1 = odense.flat_ip_search(X, Q, 17, block=700)
    D2, I2 = odense.flat_ip_search(X, Q, 17, block=1 << 20)
    assert (I1 == I2).all() and np.allclose(D1, D2, rtol=1e-6)
    D3, I3 = odense.flat_ip_search_f64(X, Q, 17)
    for i in range(9):
        check_topk(I1[i], D1[i], I3[i], D3[i], rtol=1e-5, atol=1e-6)


def test_dense_golden_from_reference_faiss_store():
    """tests/golden/dense_small.* were produced by the reference's FaissVectorStore /
    VectorStoreRetriever executed live over the numpy faiss stand-in."""
    z = np.load(os.path.join(GOLD, "dense_small.npz"))
    with open(os.path.join(GOLD, "dense_small.json")) as f:
        gold = json.load(f)
    vecs, qvecs = z["vecs"], z["qvecs"]
    for metric in ("cosine", "ip"):
        X = vecs.copy(); Qm = qvecs.copy()
        if metric == "cosine":
            odense.normalize_L2(X); odense.normalize_L2(Qm)
        for case in gold["cases"]:
            if case["kind"] != "similarity_with_score" or case["metric"] != metric:
                continue
            D, I = odense.flat_ip_search(X, Qm[case["query"]:case["query"] + 1], case["k"])
            assert I[0].tolist() == case["ids"]
            assert np.allclose(D[0], case["scores"], rtol=1e-6, atol=1e-7)


# ---- BM25 ------------------------------------------------------------------------------------------
def test_bm25_hand_computed_kat():
    corpus = [["a", "b", "a"], ["b", "c"], ["c", "c", "d", "d"]]
    bm = obm25.BM25Okapi(corpus)
    N, avgdl = 3, 9 / 3
    assert bm.avgdl == avgdl and bm.corpus_size == 3
    idf = {t: math.log(N - n + 0.5) - math.log(n + 0.5) for t, n in {"a": 1, "b": 2, "c": 2, "d": 1}.items()}
    mean = sum(idf.values()) / 4
    for t in ("b", "c"):
        assert idf[t] < 0
        idf[t] = 0.25 * mean                     # epsilon floor
    assert bm.idf == pytest.approx(idf, rel=0, abs=0)
    k1, b = 1.5, 0.75
    def term(t, tf, dl):
        return idf[t] * (tf * (k1 + 1) / (tf + k1 * (1 - b + b * dl / avgdl)))
    want = np.array([term("a", 2, 3) + term("b", 1, 3), term("b", 1, 2), 0.0])
    got = bm.get_scores(["a", "b", "zzz"])
    assert np.array_equal(got, want)
    # duplicates repeat
    assert np.array_equal(bm.get_scores(["a", "a"]), np.array([term("a", 2, 3) * 2, 0, 0]))


def test_bm25_csr_form_is_bit_identical_to_faithful_form():
    rng = np.random.default_rng(1)
    vocab = [f"w{i}" for i in range(80)]
    p = 1.0 / np.arange(1, 81); p /= p.sum()
    corpus = [rng.choice(vocab, size=int(rng.integers(1, 40)), p=p).tolist() for _ in range(400)]
    a, c = obm25.BM25Okapi(corpus), obm25.Bm25Csr(corpus)
    assert a.avgdl == c.avgdl and a.average_idf == c.average_idf
    for _ in range(25):
        q = rng.choice(vocab + ["oov"], size=int(rng.integers(1, 9))).tolist()
        assert np.array_equal(a.get_scores(q).view(np.uint64), c.get_scores(q).view(np.uint64))


def test_bm25_golden_from_reference_retriever():
    with open(os.path.join(GOLD, "bm25_hybrid_small.json")) as f:
        gold = json.load(f)
    bm = obm25.BM25Okapi([t.split() for t in gold["texts"]])
    full = {}
    for rec in gold["bm25"]:
        if "scores" in rec:
            full[rec["query"]] = np.array(rec["scores"])
            got = bm.get_scores(gold["queries"][rec["query"]].split())
            assert np.array_equal(got.view(np.uint64), full[rec["query"]].view(np.uint64))
    for rec in gold["bm25"]:
        if "ids" in rec:            # the reference's numpy argsort order: compare tie-aware
            s = full[rec["query"]]
            ref = obm25.stable_topk(s, rec["k"])
            assert np.array_equal(s[rec["ids"]], s[ref])


# ---- RRF -------------------------------------------------------------------------------------------
def test_rrf_golden_and_one_ulp_collision_cases():
    with open(os.path.join(GOLD, "rrf_reference.json")) as f:
        gold = json.load(f)
    names = set()
    for case in gold["cases"]:
        ids, scores = orrf.rrf_fuse_ids(case["lists"], case["top_k"], case["k"])
        assert ids == case["fused_ids"], case["name"]
        assert scores == case["fused_scores"], case["name"]
        assert case["fused_ranks"] == list(range(1, len(ids) + 1))
        names.add(case["name"])
    assert {"ulp_6_39_vs_12_28", "ulp_30_50_vs_39_39"} <= names
    # the two rank pairs are EQUAL as rationals but differ by one ulp in fp64 (why an exact-rational
    # comparator would mis-order against the reference)
    for (a, b), (c, d) in (((6, 39), (12, 28)), ((30, 50), (39, 39))):
        assert Fraction(1, 60 + a) + Fraction(1, 60 + b) == Fraction(1, 60 + c) + Fraction(1, 60 + d)
        assert 1.0 / (60.0 + a) + 1.0 / (60.0 + b) != 1.0 / (60.0 + c) + 1.0 / (60.0 + d)


def test_rrf_rows_oracle_returns_the_document_the_reference_returns_for_duplicated_contents():
    """tests/golden/rrf_rows_reference.json: the reference's RRFusion.fuse run live on Documents whose contents
    repeat inside one retriever's corpus and across retrievers - the Document handed back for a content is
    the last one seen in the walk (Fusion.py:61); oracle.rrf.rrf_fuse_rows must name the same (list, row)."""
    with open(os.path.join(GOLD, "rrf_rows_reference.json")) as f:
        gold = json.load(f)
    assert len(gold["cases"]) >= 16
    saw_cross_list = False
    for case in gold["cases"]:
        texts = [[f"c{x}" for x in tab] for tab in case["contents"]]
        pairs, scores = orrf.rrf_fuse_rows(case["rows"], texts, case["top_k"], case["k"])
        assert [list(p) for p in pairs] == case["fused"], case["name"]
        assert scores == case["fused_scores"], case["name"]
        first_list = {}
        for l, rws in enumerate(case["rows"]):
            for r in rws:
                first_list.setdefault(case["contents"][l][r], l)
        saw_cross_list |= any(first_list[case["contents"][l][r]] != l for l, r in pairs)
    assert saw_cross_list          # some fused Document comes from a later list than the one that introduced its content


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")
def test_rrf_oracle_matches_live_reference_on_random_lists():
    ns = ref_loader.load()
    rng = np.random.default_rng(3)
    for trial in range(40):
        L = int(rng.integers(1, 5))
        lists = [rng.permutation(120)[:int(rng.integers(0, 51))].tolist() for _ in range(L)]
        top_k = int(rng.choice([1, 7, 10, 50, 200]))
        res = [[ns.RetrievalResult(document=ns.Document(content=str(i)), score=1.0) for i in lst] for lst in lists]
        fused = ns.RRFusion().fuse(res, top_k)
        ids, scores = orrf.rrf_fuse_ids(lists, top_k)
        assert ids == [int(r.document.content) for r in fused]
        assert scores == [r.score for r in fused]


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")
def test_reference_framework_tests_still_pass_under_the_loader():
    ns = ref_loader.load()
    assert ns.Register is not None and issubclass(ns.FaissVectorStore, ns.VectorStore)


# ---- pooling ---------------------------------------------------------------------------------------
def test_pool_normalize_kat_against_torch_formulas():
    import torch
    rng = np.random.default_rng(5)
    x = rng.standard_normal((4, 6, 8)).astype(np.float32)
    mask = np.array([[1, 1, 1, 0, 0, 0], [1, 1, 1, 1, 1, 1], [1, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]])
    xt, mt = torch.from_numpy(x), torch.from_numpy(mask).float()
    mean = (xt * mt[:, :, None]).sum(1) / mt.sum(1, keepdim=True).clamp(min=1e-9)
    assert np.allclose(opool.pool_normalize(x, mask, "mean", False), mean.numpy(), rtol=1e-6, atol=1e-7)
    nrm = torch.nn.functional.normalize(mean, p=2, dim=1, eps=1e-12)
    assert np.allclose(opool.pool_normalize(x, mask, "mean", True), nrm.numpy(), rtol=1e-6, atol=1e-7)
    assert np.array_equal(opool.pool_normalize(x, mask, "cls", False), x[:, 0])
    last = opool.pool_normalize(x[:3], mask[:3], "last", False)
    assert np.array_equal(last, np.stack([x[0, 2], x[1, 5], x[2, 0]]))


# ---- comparators ---------------------------------------------------------------------------------------
def test_comparators_accept_tie_swaps_and_reject_real_errors():
    s = np.array([5.0, 4.0, 4.0, 4.0, 1.0, 0.5])
    check_topk_against_scores([0, 2, 1], [5, 4, 4], s, 3)
    check_topk_against_scores([0, 3, 2], [5, 4, 4], s, 3)            # a different member of the tie group
    with pytest.raises(AssertionError):
        check_topk_against_scores([0, 4, 1], [5, 1, 4], s, 3)        # not descending / wrong member
    with pytest.raises(AssertionError):
        check_topk_against_scores([0, 1, 1], [5, 4, 4], s, 3)        # duplicate id
    with pytest.raises(AssertionError):
        check_topk_against_scores([1, 2, 3], [4, 4, 4], s, 3)        # missed the best row
    check_topk([0, 2, 1], [5, 4, 4], [0, 1, 2], [5, 4, 4])
    with pytest.raises(AssertionError):
        check_topk([4, 1, 2], [5, 4, 4], [0, 1, 2], [5, 4, 4])        # wrong id outside any tie group


def test_adjacent_cosine_oracle_equals_live_reference_golden():
    """oracle/chunk.py against the distances the reference's own calculate_cosine_distances
    (spliter.py:354-372) produced for the committed inputs (oracle/gen_golden.py)."""
    import os
    from oracle import chunk
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "adjacent_cosine.npz"))
    got = chunk.adjacent_cosine_distances(z["emb"])
    assert np.array_equal(got, z["distances"])
    assert got[6] == 1.0 and got[7] == 1.0            # neighbours of the zero row
    assert abs(got[19]) < 1e-12                       # duplicated row
