"""GPU parity: fused dense scoring + top-k (C ABI ``ragarc_dense_topk``) against the oracle
restatement of ``faiss.IndexFlatIP.search`` (VectorStore_Faiss.py:263) on identical inputs.

Tolerances: both sides consume the same storage-dtype values, both accumulate in fp32, so scores
must agree within 1e-5 * max(1,|s|) (north_star: "1e-5 for fp32 accumulate"); ids must agree
except inside groups of scores closer than that (tie-aware comparator).
"""
import numpy as np
import pytest
import torch

from oracle import dense as odense
from oracle.compare import check_topk_against_scores
from rag_arc_b200 import _native as N
from rag_arc_b200 import ops, synth

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-5


def _check_all(ids, scores, X, Q, k, what):
    S = Q.astype(np.float64) @ X.astype(np.float64).T
    ids = ids.cpu().numpy(); scores = scores.cpu().numpy()
    for i in range(Q.shape[0]):
        check_topk_against_scores(ids[i], scores[i], S[i], k, rtol=RTOL, atol=ATOL, what=f"{what} q{i}")


def test_c1_fp32_flat_ip_matches_faiss_restatement(dev):
    """BASELINE config 1: 10k x 384 normalised fp32, 100 queries, top-10."""
    X = synth.dense_corpus_np(10_000, 384)
    Q, planted = synth.dense_queries_np(X, 100)
    D, I = odense.flat_ip_search(X, Q, 10)
    scores, ids, path = ops.dense_topk(torch.from_numpy(X).to(dev), torch.from_numpy(Q).to(dev), 10,
                                       return_path=True)
    assert path == N.DENSE_SIMT
    assert (ids.cpu().numpy() == I).all()
    assert np.allclose(scores.cpu().numpy(), D, rtol=RTOL, atol=ATOL)
    assert (ids[:, 0].cpu().numpy() == planted).all()
    _check_all(ids, scores, X, Q, 10, "c1")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,d,nq,k", [(5000, 64, 7, 5), (33_000, 768, 130, 100), (70_001, 384, 256, 50),
                                      (1000, 1024, 1, 10), (257, 72, 3, 200)])
def test_tcgen05_path_matches_oracle(dev, dtype, n, d, nq, k):
    x = synth.dense_corpus_cuda(n, d, dtype, dev, seed=7)
    q, planted = synth.dense_queries_cuda(x, nq, seed=11)
    scores, ids, path = ops.dense_topk(x, q, k, return_path=True)
    assert path == N.DENSE_TCGEN05
    X = x.float().cpu().numpy(); Q = q.float().cpu().numpy()
    _check_all(ids, scores, X, Q, k, f"tc {dtype} {n}x{d}")
    assert (ids[:, 0].cpu() == planted.cpu()).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_simt_and_tc_agree_bitwise_on_ids(dev, dtype):
    """Same bf16 inputs through both kernels: identical ids, scores within fp32 summation noise."""
    n, d, nq, k = 20_000, 256, 64, 20
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev, seed=3).to(dtype)
    q, _ = synth.dense_queries_cuda(x, nq, seed=5)
    s1, i1 = ops.dense_topk(x, q, k, path=N.DENSE_SIMT)
    if dtype == torch.float32:
        _check_all(i1, s1, x.cpu().numpy(), q.cpu().numpy(), k, "simt f32")
        return
    s2, i2 = ops.dense_topk(x, q, k, path=N.DENSE_TCGEN05)
    assert torch.allclose(s1, s2, rtol=RTOL, atol=ATOL)
    X = x.float().cpu().numpy(); Q = q.float().cpu().numpy()
    _check_all(i1, s1, X, Q, k, "simt bf16")
    _check_all(i2, s2, X, Q, k, "tc bf16")


def test_k_larger_than_corpus_pads_with_minus_one(dev):
    x = synth.dense_corpus_cuda(37, 64, torch.bfloat16, dev)
    q, _ = synth.dense_queries_cuda(x, 4)
    for path in (N.DENSE_SIMT, N.DENSE_TCGEN05):
        scores, ids = ops.dense_topk(x, q, 50, path=path)
        ids = ids.cpu().numpy(); scores = scores.cpu().numpy()
        assert (ids[:, 37:] == -1).all() and np.isinf(scores[:, 37:]).all()
        for i in range(4):
            assert sorted(ids[i, :37].tolist()) == list(range(37))
            assert (np.diff(scores[i, :37]) <= 0).all()


def test_ties_resolve_to_lowest_row_id(dev):
    """Duplicate rows give exactly equal scores; the lowest row id must come first, and a zero
    query (all scores equal) must return rows 0..k-1."""
    base = synth.dense_corpus_cuda(500, 128, torch.bfloat16, dev, seed=1)
    x = torch.cat([base, base, base], dim=0).contiguous()        # rows r, r+500, r+1000 identical
    q, planted = synth.dense_queries_cuda(base, 9, seed=2)
    q = torch.cat([q, torch.zeros((1, 128), dtype=q.dtype, device=dev)], dim=0).contiguous()
    for path in (N.DENSE_SIMT, N.DENSE_TCGEN05):
        scores, ids = ops.dense_topk(x, q, 6, path=path)
        ids = ids.cpu().numpy(); scores = scores.cpu().numpy()
        for i in range(9):
            p = int(planted[i])
            assert ids[i, :3].tolist() == [p, p + 500, p + 1000]
            assert scores[i, 0] == scores[i, 1] == scores[i, 2]
            r = ids[i, 3]
            assert ids[i, 3:6].tolist() == [r, r + 500, r + 1000]
        assert ids[9].tolist() == [0, 1, 2, 3, 4, 5] and (scores[9] == 0).all()


def test_ascending_scores_worst_case_for_running_threshold(dev):
    """Rows ordered so that every later row scores higher: the running threshold never helps and
    every list must be pruned again and again; result must still be exact."""
    n, d, k = 6000, 64, 10
    x = torch.zeros((n, d), dtype=torch.float32)
    x[:, 0] = torch.linspace(0.001, 1.0, n)
    x[:, 1] = 0.25
    x = x.to(torch.bfloat16).to(dev)
    q = torch.zeros((3, d), dtype=torch.bfloat16, device=dev)
    q[:, 0] = 1.0
    for path in (N.DENSE_SIMT, N.DENSE_TCGEN05):
        scores, ids = ops.dense_topk(x, q, k, path=path)
        _check_all(ids, scores, x.float().cpu().numpy(), q.float().cpu().numpy(), k, "ascending")


def test_rows_subset_and_result_independent_of_it(dev):
    """``n_rows`` < allocated rows (the store keeps spare capacity): rows past it are ignored."""
    x = synth.dense_corpus_cuda(4096, 128, torch.bfloat16, dev, seed=9)
    q, _ = synth.dense_queries_cuda(x, 5, seed=10, n_rows=1000)
    s_a, i_a = ops.dense_topk(x, q, 8, n_rows=1000)
    s_b, i_b = ops.dense_topk(x[:1000].contiguous(), q, 8)
    assert torch.equal(i_a, i_b) and torch.equal(s_a, s_b)
    assert int(i_a.max()) < 1000


def test_keys_and_merge_equal_single_shot(dev):
    """Shard the corpus in 3 uneven pieces, search each, merge the packed keys: must be bitwise the
    single-shot result (the multi-GPU path's arithmetic, on one device)."""
    n, d, nq, k = 50_000, 128, 33, 20
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev, seed=21)
    q, _ = synth.dense_queries_cuda(x, nq, seed=22)
    s_ref, i_ref = ops.dense_topk(x, q, k)
    cuts = [0, 17_000, 17_100, n]
    keys = [ops.dense_topk_keys(x[a:b].contiguous(), q, k, id_base=a) for a, b in zip(cuts[:-1], cuts[1:])]
    s, i = ops.merge_topk_keys(torch.stack(keys, 0).contiguous(), k)
    assert torch.equal(i, i_ref) and torch.equal(s, s_ref)


@pytest.mark.parametrize("signal", [False, True])
@pytest.mark.parametrize("n,nq,k,G", [(50_000, 33, 20, 3), (200_000, 520, 100, 4), (1_000, 5, 10, 8)])
def test_query_owner_push_exchange_equals_single_shot(dev, n, nq, k, G, signal):
    """The multi-GPU query-owner exchange on one device: G row shards, every shard's merge kernel
    pushes the key row of query q into the inbox of rank q // nq_per (here: G inboxes in local
    memory), every owner merges its own queries - concatenated, the owners' results must be bitwise
    the single-shot result.  signal=True: the pushers count themselves into per-query arrival counters
    behind each inbox and the owner's merge kernel waits on them (no barrier between the kernels);
    run twice to check that the counters are reset."""
    d = 128
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev, seed=21)
    q, _ = synth.dense_queries_cuda(x, nq, seed=22)
    s_ref, i_ref = ops.dense_topk(x, q, k)
    nq_per = (nq + G - 1) // G
    words = ops.inbox_words(G, nq_per, k)
    inbox = torch.zeros((G, words), dtype=torch.int64, device=dev)             # [owner][keys | counters]
    table = torch.tensor([inbox[o].data_ptr() for o in range(G)], dtype=torch.int64, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    per = (n + G - 1) // G
    for rep in range(2 if signal else 1):
        for g in range(G):
            lo, hi = min(n, g * per), min(n, (g + 1) * per)
            ops.dense_topk_keys_push(x[lo:hi].contiguous(), q, k, lo, table, g, nq_per, signal=signal)
        outs_s, outs_i = [], []
        for o in range(G):
            n_own = max(0, min(nq, (o + 1) * nq_per) - o * nq_per)
            if signal:
                s, i = ops.merge_topk_inbox(inbox[o], G, nq_per, n_own, k, k, status, timeout_ms=200.0)
            else:
                s, i = ops.merge_topk_keys(inbox[o][:G * nq_per * k].view(G, nq_per, k), k)
            outs_s.append(s[:n_own]); outs_i.append(i[:n_own])
        assert torch.equal(torch.cat(outs_i), i_ref) and torch.equal(torch.cat(outs_s), s_ref)
        assert int(status.item()) == 0
        if signal:      # every counter was consumed and reset
            cnt = inbox[:, G * nq_per * k:].view(torch.int32)
            assert int(cnt.abs().sum()) == 0


def test_published_rung_thresholds_do_not_change_results(dev):
    """The running threshold built from published order statistics (replaces the seeding pass)
    is only a filter: a shape that uses it must still match the fp64 oracle exactly, including a
    heavily duplicated corpus where many rows tie with the published value."""
    n, d, nq, k = 120_000, 128, 300, 100
    plan = N.dense_plan(n, d, N.BF16, nq, k)
    assert plan["publishing_lists"] > 0 and plan["seed_rows"] == 0, plan
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev, seed=77)
    q, planted = synth.dense_queries_cuda(x, nq, seed=78)
    scores, ids = ops.dense_topk(x, q, k)
    _check_all(ids, scores, x.float().cpu().numpy(), q.float().cpu().numpy(), k, "rungs")
    assert (ids[:, 0] == planted).all()
    base = x[:3000]
    xd = base.repeat(40, 1).contiguous()                      # every row 40 times: 40-way exact ties
    s2, i2 = ops.dense_topk(xd, q[:64].contiguous(), k)
    S = q[:64].float() @ base.float().T
    want_s, want_r = torch.sort(S, dim=1, descending=True, stable=True)
    # top-100 of the tiled corpus = best 3 distinct rows x 40 copies (lowest ids first inside a tie)
    for qi in range(64):
        got = i2[qi].tolist()
        assert sorted(got[:40]) == got[:40] and all(g % 3000 == got[0] % 3000 for g in got[:40])
        assert got[0] % 3000 == int(want_r[qi, 0])
        assert got[:40] == [int(want_r[qi, 0]) + 3000 * j for j in range(40)]


def test_massive_ties_take_the_merge_spill_path(dev):
    """200k identical rows: every candidate ties with every threshold, nothing can be filtered, every
    list is cut to its budget by the exact prune and the merge kernel gets more survivors than its
    shared-memory array holds (global spill row).  The answer is fully determined by the tie rule:
    rows 0..k-1 in order, for every query, on both kernels' shared merge."""
    d, k = 64, 100
    row = torch.nn.functional.normalize(torch.randn((1, d), generator=torch.Generator().manual_seed(5)), dim=1)
    x = row.repeat(200_000, 1).to(torch.bfloat16).to(dev).contiguous()
    q = torch.cat([row, -row, torch.randn((5, d), generator=torch.Generator().manual_seed(6))], 0).to(torch.bfloat16).to(dev)
    for nq in (7, 300):
        qq = q.repeat((nq + 6) // 7, 1)[:nq].contiguous()
        scores, ids = ops.dense_topk(x, qq, k)
        want = torch.arange(k, device=dev)[None, :].expand(nq, k)
        assert torch.equal(ids, want)
        assert (scores == scores[:, :1]).all()
    # a planted better row at the very end must still come first
    x[-1] = (row * 2).to(torch.bfloat16).to(dev)[0]
    scores, ids = ops.dense_topk(x, q[:1].contiguous(), k)
    assert ids[0, 0].item() == 199_999 and ids[0, 1:].tolist() == list(range(k - 1))


def test_full_size_properties_1m_x_768(dev):
    """BASELINE config 3 at full size (1M x 768 bf16, 1024 queries, top-100) through
    size-independent properties: planted neighbour at rank 1, scores descending, ids unique and
    in range, every returned score equals the recomputed fp32 dot product, and nothing in a
    random 20k-row sample beats the k-th score."""
    n, d, nq, k = 1_000_000, 768, 1024, 100
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev)
    q, planted = synth.dense_queries_cuda(x, nq)
    scores, ids, path = ops.dense_topk(x, q, k, return_path=True)
    assert path == N.DENSE_TCGEN05
    assert (ids[:, 0] == planted).all()
    assert (scores[:, 1:] <= scores[:, :-1]).all()
    assert int(ids.min()) >= 0 and int(ids.max()) < n
    srt = ids.sort(dim=1).values
    assert (srt[:, 1:] != srt[:, :-1]).all()
    for i in range(0, nq, 64):
        rec = (x[ids[i]].float() @ q[i].float())
        assert torch.allclose(rec, scores[i], rtol=1e-5, atol=1e-5)
    g = torch.Generator(device=dev); g.manual_seed(99)
    sample = torch.randint(0, n, (20_000,), generator=g, device=dev)
    S = q.float() @ x[sample].float().T                    # [nq, 20000]
    kth = scores[:, -1:]
    beats = S > kth + 1e-5
    # any sampled row that beats the k-th score must already be in the result
    rows_q, cols = beats.nonzero(as_tuple=True)
    for rq, c in zip(rows_q.tolist()[:2000], cols.tolist()[:2000]):
        assert (ids[rq] == sample[c]).any()


@pytest.mark.parametrize("n,d,nq,k", [(10_000, 384, 100, 10), (60_000, 768, 300, 100), (3000, 64, 5, 7)])
def test_bf16x3_tensor_core_path_is_fp32_accurate(dev, n, d, nq, k):
    """fp32 corpus, fp32 queries, searched on the tensor cores through three bf16 planes: scores
    must match the fp32 oracle (FAISS restatement) to 1e-5 and ids exactly (tie-aware)."""
    X = synth.dense_corpus_np(n, d, seed=31)
    Q, planted = synth.dense_queries_np(X, nq, seed=32)
    xp = ops.normalize_split3(torch.from_numpy(X).to(dev), normalize=False)
    qp = ops.normalize_split3(torch.from_numpy(Q).to(dev), normalize=False)
    # the planes reconstruct the fp32 values to 2^-24 relative
    rec = xp.float().view(n, 3, d).sum(1).cpu().numpy()
    assert np.abs(rec - X).max() <= 2.0 ** -22 * np.abs(X).max()
    scores, ids = ops.dense_topk_x3(xp, qp, k)
    _check_all(ids, scores, X, Q, k, "bf16x3")
    assert (ids[:, 0].cpu().numpy() == planted).all()
    D, I = odense.flat_ip_search(X, Q, k)
    assert np.allclose(scores.cpu().numpy(), D, rtol=1e-5, atol=1e-6)
    assert (ids.cpu().numpy() == I).mean() > 0.999


def test_concurrent_host_threads_on_their_own_streams(dev):
    """The ABI may be entered from arbitrary host threads (the reference runs retrievers in a thread
    pool, core/retrieval/base.py:82-96): four threads, each on its own CUDA stream, search the same
    corpus concurrently (512 queries -> the multicast-cluster schedule with its side-stream launch)."""
    from concurrent.futures import ThreadPoolExecutor
    n, d, k = 150_000, 128, 20
    x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev, seed=41)
    qs = [synth.dense_queries_cuda(x, 512, seed=50 + i)[0] for i in range(4)]
    want = [ops.dense_topk(x, q, k) for q in qs]
    torch.cuda.synchronize()

    def work(i):
        st = torch.cuda.Stream(dev)
        outs = []
        with torch.cuda.stream(st):
            for _ in range(6):
                outs.append(ops.dense_topk(x, qs[i], k))
        st.synchronize()
        return all(torch.equal(s, want[i][0]) and torch.equal(r, want[i][1]) for s, r in outs)

    with ThreadPoolExecutor(4) as ex:
        assert all(ex.map(work, range(4)))
