"""GPU: seeded random shapes through ``ragarc_dense_topk`` (both kernels), ``ragarc_dense_topk_keys``
+ merge, and ``ragarc_bm25_topk`` against the oracle - ragged sizes (rows, queries and k that are no
multiple of any tile), every list-capacity class of k (<= 224 / 480 / 992 / 2016), tiny corpora, all
three storage dtypes.  Complements the fixed-shape cases of test_gpu_dense.py."""
import numpy as np
import pytest
import torch

from oracle import bm25 as obm25
from oracle.compare import check_topk_against_scores
from rag_arc_b200 import _native as N
from rag_arc_b200 import ops
from rag_arc_b200.core.retrieval.bm25_index import Bm25Index

pytestmark = pytest.mark.gpu


def _cases(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        n = int(rng.choice([1, 2, 31, 255, 256, 257, 1000, 4097, 20_011, 66_000]))
        d = int(rng.choice([8, 24, 64, 72, 200, 384, 776, 1024]))
        nq = int(rng.choice([1, 2, 17, 127, 128, 129, 255, 300, 513]))
        k = int(rng.choice([1, 3, 10, 100, 224, 225, 480, 481, 992, 1500, 2016]))
        dt = [torch.bfloat16, torch.float16, torch.float32][int(rng.integers(0, 3))]
        out.append((n, d, nq, k, dt, int(rng.integers(0, 1 << 30))))
    return out


@pytest.mark.parametrize("n,d,nq,k,dtype,seed", _cases(2024, 28))
def test_random_shape_dense_topk_matches_oracle(dev, n, d, nq, k, dtype, seed):
    if nq * n * (1 if k <= 224 else 4) > 40_000_000:
        nq = max(1, 40_000_000 // (n * 4))                      # bound the host-side oracle work
    g = torch.Generator().manual_seed(seed)
    X = torch.nn.functional.normalize(torch.randn((n, d), generator=g), dim=1).to(dtype)
    Q = torch.nn.functional.normalize(torch.randn((nq, d), generator=g), dim=1).to(dtype)
    S = Q.double().numpy() @ X.double().numpy().T
    x, q = X.to(dev), Q.to(dev)
    paths = [N.DENSE_SIMT] if dtype == torch.float32 else [N.DENSE_SIMT, N.DENSE_TCGEN05]
    for path in paths:
        scores, ids = ops.dense_topk(x, q, k, path=path)
        scores = scores.cpu().numpy(); ids = ids.cpu().numpy()
        kk = min(k, n)
        assert (ids[:, kk:] == -1).all()
        for i in range(nq):
            check_topk_against_scores(ids[i, :kk], scores[i, :kk], S[i], kk, rtol=1e-5, atol=1e-5,
                                      what=f"path{path} n={n} d={d} nq={nq} k={k} {dtype} q{i}")
    # sharded form on one device: two uneven shards, packed keys, merge == single shot (bitwise)
    if n >= 2:
        cut = max(1, n // 3)
        s_ref, i_ref = ops.dense_topk(x, q, k)
        keys = [ops.dense_topk_keys(x[a:b].contiguous(), q, k, id_base=a) for a, b in ((0, cut), (cut, n))]
        s, i = ops.merge_topk_keys(torch.stack(keys, 0).contiguous(), k)
        assert torch.equal(i, i_ref) and torch.equal(s, s_ref)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_bm25_corpora_bit_exact(dev, seed):
    rng = np.random.default_rng(seed)
    n_docs = int(rng.choice([1, 5, 300, 24_576, 24_577, 60_000]))
    vocab = int(rng.choice([3, 50, 2000]))
    words = [f"w{i}" for i in range(vocab)]
    lens = rng.integers(0 if n_docs > 5 else 1, 30, n_docs)     # empty documents allowed
    if lens.sum() == 0:
        lens[0] = 3
    corpus = [[words[j] for j in rng.zipf(1.3, L) % vocab] for L in lens]
    queries = [[words[j] for j in rng.integers(0, vocab, int(rng.integers(1, 12)))] + (["oov"] if qi % 3 == 0 else [])
               for qi in range(9)]
    queries.append([queries[0][0]] * 4)                           # repeated term counts four times
    k = int(rng.choice([1, 7, 64, 1000]))
    idx = Bm25Index.from_token_lists(corpus, device=dev)
    qt, ql = idx.encode_queries(queries)
    sc, ids = ops.bm25_topk(idx, qt, ql, k)
    sc = sc.cpu().numpy(); ids = ids.cpu().numpy()
    ref = obm25.BM25Okapi(corpus)
    kk = min(k, n_docs)
    for qi, qtok in enumerate(queries):
        want = ref.get_scores(qtok)
        top = obm25.stable_topk(want, kk)
        assert ids[qi, :kk].tolist() == top.tolist(), f"seed {seed} q{qi}"
        assert np.array_equal(sc[qi, :kk].view(np.uint64), want[top].view(np.uint64))
        assert (ids[qi, kk:] == -1).all()
