"""GPU parity for BM25 (bit-exact fp64), RRF (bit-exact ranking + fp64 scores) and pool+normalise,
each through the C ABI against the oracle on identical inputs."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bm25 as obm25
from oracle import pool as opool
from oracle import rrf as orrf
from rag_arc_b200 import ops, synth
from rag_arc_b200.core.retrieval.bm25_index import Bm25Index

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tie_aware_ids(ids, scores_full, k):
    """ids must be a valid top-k of scores_full: same score multiset as the stable top-k and every
    id carries exactly the score at its position."""
    ref = obm25.stable_topk(scores_full, k)
    assert np.array_equal(scores_full[ids], scores_full[ref])
    assert len(set(ids.tolist())) == len(ids)


def test_bm25_small_text_corpus_bit_exact(dev):
    texts = ["the quick brown fox jumps over the lazy dog", "the lazy dog sleeps", "quick quick fox",
             "a completely different sentence about cats", "dog dog dog the the", "fox"]
    toks = [t.split() for t in texts]
    ref = obm25.BM25Okapi(toks)
    idx = Bm25Index.from_token_lists(toks, device=dev)
    queries = [["quick", "fox"], ["the", "dog", "the"], ["unseen", "cats"], ["zzz"], []]
    qt, ql = idx.encode_queries(queries)
    full = ops.bm25_scores(idx, qt, ql).cpu().numpy()
    sc, ids = ops.bm25_topk(idx, qt, ql, 4)
    sc = sc.cpu().numpy(); ids = ids.cpu().numpy()
    for i, q in enumerate(queries):
        want = ref.get_scores(q)
        assert np.array_equal(full[i].view(np.uint64), want.view(np.uint64)), f"q{i} scores not bit-exact"
        _tie_aware_ids(ids[i], want, 4)
        assert np.array_equal(sc[i].view(np.uint64), want[ids[i]].view(np.uint64))
        # our documented tie rule: descending score, ascending id
        assert ids[i].tolist() == obm25.stable_topk(want, 4).tolist()


def test_bm25_c2_shape_bit_exact_against_csr_oracle(dev):
    """BASELINE config 2 sparse side: 100k docs, Zipf vocabulary, 256 queries x 8 tokens, top-50."""
    toks, offs = synth.bm25_corpus_tokens(100_000)
    qtok = synth.bm25_queries_tokens(toks, offs, 256)
    idx = Bm25Index.from_token_ids(toks, offs, device=dev)
    qt, ql = idx.encode_query_ids(qtok)
    sc, ids = ops.bm25_topk(idx, qt, ql, 50)
    sc = sc.cpu().numpy(); ids = ids.cpu().numpy()
    # oracle over the same postings, integer tokens (oracle/bm25.py Bm25Csr is checked against the
    # faithful dict-based restatement in tests/test_oracle.py)
    ref = obm25.Bm25Csr([toks[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)])
    for i in range(0, 256, 5):
        want = ref.get_scores(qtok[i].tolist())
        _tie_aware_ids(ids[i], want, 50)
        assert np.array_equal(sc[i].view(np.uint64), want[ids[i]].view(np.uint64)), f"q{i}"
        assert ids[i].tolist() == obm25.stable_topk(want, 50).tolist()
    # the in-kernel evaluation of the per-posting factor (no precomputed table) is bit-identical
    sc2, ids2 = ops.bm25_topk(idx, qt, ql, 50, use_post_val=False)
    assert np.array_equal(sc2.cpu().numpy().view(np.uint64), sc.view(np.uint64))
    assert np.array_equal(ids2.cpu().numpy(), ids)
    full_a = ops.bm25_scores(idx, qt[:8].contiguous(), ql[:8].contiguous())
    full_b = ops.bm25_scores(idx, qt[:8].contiguous(), ql[:8].contiguous(), use_post_val=False)
    assert torch.equal(full_a.view(torch.int64), full_b.view(torch.int64))


def test_bm25_doc_range_shards_merge_to_the_single_index_result(dev):
    """Three uneven doc-range shards (global idf / avgdl), per-shard top-k, ragarc_bm25_merge_topk:
    bit-identical to the unsharded search - the multi-GPU arithmetic on one device."""
    toks, offs = synth.bm25_corpus_tokens(30_000, vocab=6000, seed=21)
    full = Bm25Index.from_token_ids(toks, offs, device=None)
    single = Bm25Index.from_token_ids(toks, offs, device=dev)
    qt, ql = single.encode_query_ids(synth.bm25_queries_tokens(toks, offs, 40, 8, seed=22))
    for k in (20, 700):
        s_ref, i_ref = ops.bm25_topk(single, qt, ql, k)
        parts = []
        for lo, hi in ((0, 11_000), (11_000, 11_300), (11_300, 30_000)):
            sh = full.shard(lo, hi, dev)
            kk = min(k, hi - lo)
            s, i = ops.bm25_topk(sh, qt, ql, kk)
            ps = torch.full((40, k), float("-inf"), dtype=torch.float64, device=dev)
            pi = torch.full((40, k), -1, dtype=torch.int64, device=dev)
            ps[:, :kk] = s; pi[:, :kk] = i
            parts.append((ps, pi))
        s, i = ops.bm25_merge_topk(torch.stack([p[0] for p in parts]).contiguous(),
                                   torch.stack([p[1] for p in parts]).contiguous(), k)
        assert torch.equal(i, i_ref)
        assert torch.equal(s.view(torch.int64), s_ref.view(torch.int64))


def test_bm25_fewer_matches_than_k_fills_with_zero_score_docs(dev):
    toks = [["a", "b"], ["c"], ["d"], ["e"], ["a"], ["f"]]
    idx = Bm25Index.from_token_lists(toks, device=dev)
    qt, ql = idx.encode_queries([["a"]])
    sc, ids = ops.bm25_topk(idx, qt, ql, 5)
    want = obm25.BM25Okapi(toks).get_scores(["a"])
    assert ids[0].tolist() == obm25.stable_topk(want, 5).tolist()
    assert np.array_equal(sc[0].cpu().numpy(), want[ids[0].cpu().numpy()])
    sc, ids = ops.bm25_topk(idx, qt, ql, 10)        # k > n_docs -> padded
    assert ids[0, 6:].tolist() == [-1] * 4


def test_rrf_matches_integer_oracle_and_reference_golden(dev):
    rng = np.random.default_rng(5)
    L, nq, kl, top_k = 2, 64, 50, 10
    ids = np.stack([np.stack([rng.permutation(400)[:kl] for _ in range(nq)]) for _ in range(L)]).astype(np.int32)
    ids[1, 3, 40:] = -1                     # a short list
    ids[0, 4, :] = -1                       # an empty list (failed retriever)
    out_ids, out_sc, out_n = ops.rrf_fuse(torch.from_numpy(ids).to(dev), top_k)
    out_ids = out_ids.cpu().numpy(); out_sc = out_sc.cpu().numpy(); out_n = out_n.cpu().numpy()
    for q in range(nq):
        want_ids, want_sc = orrf.rrf_fuse_ids([ids[l, q].tolist() for l in range(L)], top_k)
        n = len(want_ids)
        assert out_n[q] == n
        assert out_ids[q, :n].tolist() == want_ids
        assert np.array_equal(out_sc[q, :n].view(np.uint64), np.array(want_sc).view(np.uint64))
    # golden vectors produced by the reference's own RRFusion (oracle/gen_golden.py)
    with open(os.path.join(GOLD, "rrf_reference.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        lists = case["lists"]
        kl = max(1, max(len(l) for l in lists))
        arr = np.full((len(lists), 1, kl), -1, np.int32)
        for l, lst in enumerate(lists):
            arr[l, 0, :len(lst)] = lst
        o_ids, o_sc, o_n = ops.rrf_fuse(torch.from_numpy(arr).to(dev), case["top_k"], case["k"])
        n = int(o_n[0])
        assert o_ids[0, :n].tolist() == case["fused_ids"], case["name"]
        got = o_sc[0, :n].cpu().numpy()
        assert np.array_equal(got.view(np.uint64), np.array(case["fused_scores"], np.float64).view(np.uint64)), case["name"]


def test_rrf_fuse_rows_resolves_documents_like_the_reference_walk(dev):
    """ragarc_rrf_fuse_rows against oracle.rrf.rrf_fuse_rows (the reference's walk over Documents stated on
    rows): three retrievers over corpora that share contents (within one corpus and across corpora), lists
    of different widths, trailing padding, an all-padding list and a retriever that returned nothing; scores
    bit-exact, (list, row) exact, and the key-level output equal to ragarc_rrf_fuse on the same keys."""
    rng = np.random.default_rng(11)
    nq, kl, top_k = 97, 50, 10
    sizes, widths = [300, 200, 120], [50, 37, 50]
    n_contents = 260                                                 # < total rows: many duplicates
    contents = [rng.integers(0, n_contents, n).astype(np.int32) for n in sizes]      # row -> content key
    rows = [np.stack([rng.permutation(n)[:w] for _ in range(nq)]).astype(np.int64) for n, w in zip(sizes, widths)]
    rows[0][5, 30:] = -1
    rows[1][6, :] = -1
    tabs = [torch.from_numpy(c).to(dev) for c in contents]
    for variant in ("all", "middle-missing"):
        dev_rows = [torch.from_numpy(r).to(dev) for r in rows]
        host_rows = [r for r in rows]
        if variant == "middle-missing":
            dev_rows[1] = None
        out_ids, out_sc, packed = ops.rrf_fuse_rows(dev_rows, tabs, kl, top_k, 60.0)
        lists, rws, cnt = ops.unpack_fused_rows(packed.cpu().numpy(), nq, top_k)
        out_ids = out_ids.cpu().numpy(); out_sc = out_sc.cpu().numpy()
        for q in range(nq):
            per = [([] if dev_rows[l] is None else host_rows[l][q].tolist()) for l in range(3)]
            pairs, scores = orrf.rrf_fuse_rows(per, [c.tolist() for c in contents], top_k, 60.0)
            n = len(pairs)
            assert cnt[q] == n
            assert list(zip(lists[q, :n].tolist(), rws[q, :n].tolist())) == pairs, (variant, q)
            assert np.array_equal(out_sc[q, :n].view(np.uint64), np.array(scores).view(np.uint64))
            assert out_ids[q, :n].tolist() == [int(contents[l][r]) for l, r in pairs]
            assert (lists[q, n:] == -1).all() and (rws[q, n:] == -1).all()
        # same fusion as the key-level entry point on the padded key matrix
        keys = np.full((3, nq, kl), -1, np.int32)
        for l in range(3):
            if dev_rows[l] is not None:
                r = host_rows[l]
                keys[l, :, :r.shape[1]] = np.where(r >= 0, contents[l][np.clip(r, 0, None)], -1)
        k_ids, k_sc, k_n = ops.rrf_fuse(torch.from_numpy(keys).to(dev), top_k)
        assert np.array_equal(k_ids.cpu().numpy(), out_ids) and np.array_equal(k_n.cpu().numpy(), cnt)
        assert np.array_equal(k_sc.cpu().numpy().view(np.uint64), out_sc.view(np.uint64))
    # vectors produced by the reference's own RRFusion on Documents with duplicated contents
    with open(os.path.join(GOLD, "rrf_rows_reference.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        L = len(case["rows"])
        kl_c = max(1, max(len(r) for r in case["rows"]))
        d_rows = [torch.tensor([r], dtype=torch.int64, device=dev).reshape(1, len(r)) if len(r) else None for r in case["rows"]]
        d_tabs = [torch.tensor(t, dtype=torch.int32, device=dev) for t in case["contents"]]
        if all(r is None for r in d_rows):
            assert case["fused"] == []
            continue
        _, o_sc, packed = ops.rrf_fuse_rows(d_rows, d_tabs, kl_c, case["top_k"], case["k"])
        lists_, rws_, cnt_ = ops.unpack_fused_rows(packed.cpu().numpy(), 1, case["top_k"])
        n = int(cnt_[0])
        assert [[int(a), int(b)] for a, b in zip(lists_[0, :n], rws_[0, :n])] == case["fused"], case["name"]
        assert np.array_equal(o_sc[0, :n].cpu().numpy().view(np.uint64),
                              np.array(case["fused_scores"], np.float64).view(np.uint64)), case["name"]
    with pytest.raises(Exception):
        ops.rrf_fuse_rows([None, None, None], tabs, kl, top_k)
    with pytest.raises(Exception):
        ops.rrf_fuse_rows([torch.from_numpy(rows[0]).to(dev)] * 9, tabs * 3, kl, top_k)      # more than 8 lists


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("mode", ["mean", "cls", "last"])
@pytest.mark.parametrize("normalize", [True, False])
def test_pool_normalize_matches_oracle(dev, dtype, mode, normalize):
    rng = np.random.default_rng(3)
    B, T, H = 9, 37, 384
    x = torch.from_numpy(rng.standard_normal((B, T, H)).astype(np.float32)).to(dtype)
    lens = rng.integers(1, T + 1, size=B)
    mask = (np.arange(T)[None, :] < lens[:, None]).astype(np.int64)
    want = opool.pool_normalize(x.float().numpy(), mask, mode, normalize)
    got = ops.pool_normalize(x.to(dev), torch.from_numpy(mask).to(dev), mode, normalize).cpu().numpy()
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_pool_all_masked_row_and_left_padding(dev):
    x = torch.randn((2, 5, 64), device=dev)
    mask = torch.tensor([[0, 0, 0, 0, 0], [0, 0, 1, 1, 1]], device=dev)
    got = ops.pool_normalize(x, mask, "mean", False)
    assert torch.equal(got[0], torch.zeros(64, device=dev))        # 0 / clamp(0, 1e-9)
    assert torch.allclose(got[1], x[1, 2:].mean(0), rtol=1e-5, atol=1e-6)
    last = ops.pool_normalize(x[1:], mask[1:], "last", False)
    assert torch.equal(last[0], x[1, 4])
