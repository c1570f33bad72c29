"""GPU: the library-owned flat index (``ragarc_index_*``) driven through ctypes with numpy HOST
buffers only - the calls FaissVectorStore makes on faiss.IndexFlatIP (VectorStore_Faiss.py:114-115,
178, 202, 259-263, 385-419) - against the oracle restatement of that index."""
import numpy as np
import pytest

from oracle import dense as odense
from oracle.compare import check_topk_against_scores
from rag_arc_b200 import synth
from rag_arc_b200.native_index import NativeFlatIndex

pytestmark = pytest.mark.gpu


def _round_bf16(a):
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32)


def test_fp32_cosine_index_matches_faiss_restatement_c1(dev):
    """BASELINE config 1 through the handle API: un-normalised fp32 rows in, cosine metric."""
    rng = np.random.default_rng(7)
    X = rng.standard_normal((10_000, 384)).astype(np.float32) * 3.0
    Q = (X[rng.integers(0, 10_000, 100)] + 0.3 * rng.standard_normal((100, 384))).astype(np.float32)
    ix = NativeFlatIndex(384, "float32", "cosine")
    ix.add(X[:3000]); ix.add(X[3000:3001]); ix.add(X[3001:])          # grows twice
    assert ix.ntotal == 10_000
    D, I = ix.search(Q, 10)
    Xn, Qn = X.copy(), Q.copy()
    odense.normalize_L2(Xn); odense.normalize_L2(Qn)          # in place, like faiss.normalize_L2
    Dref, Iref = odense.flat_ip_search(Xn, Qn, 10)
    assert (I == Iref).all()
    assert np.allclose(D, Dref, rtol=1e-5, atol=1e-5)
    ix.close()


def test_bf16_ip_index_and_k_larger_than_ntotal(dev):
    X = synth.dense_corpus_np(5000, 256, seed=3)
    Q, planted = synth.dense_queries_np(X, 33, seed=4)
    ix = NativeFlatIndex(256, "bfloat16", "ip")
    ix.add(X)
    D, I = ix.search(Q, 20)
    S = _round_bf16(Q).astype(np.float64) @ _round_bf16(X).astype(np.float64).T
    for i in range(33):
        check_topk_against_scores(I[i], D[i], S[i], 20, rtol=1e-5, atol=1e-5, what=f"q{i}")
    assert (I[:, 0] == planted).all()
    small = NativeFlatIndex(256, "bfloat16", "ip")
    small.add(X[:7])
    D, I = small.search(Q[:2], 12)
    assert (I[:, 7:] == -1).all() and np.isinf(D[:, 7:]).all()
    assert sorted(I[0, :7].tolist()) == list(range(7))


def test_remove_renumbers_like_faiss_remove_ids(dev):
    X = synth.dense_corpus_np(2000, 128, seed=9)
    Q, _ = synth.dense_queries_np(X, 16, seed=10)
    ix = NativeFlatIndex(128, "float32", "ip")
    ix.add(X)
    drop = np.array([5, 0, 1999, 700, 5, 701])
    assert ix.remove(drop) == 5
    keep = np.setdiff1d(np.arange(2000), drop)
    D, I = ix.search(Q, 8)
    Dref, Iref = odense.flat_ip_search(X[keep], Q, 8)
    assert (I == Iref).all() and np.allclose(D, Dref, rtol=1e-5, atol=1e-5)
    ix.add(X[:3])                                                      # appended after the survivors
    assert ix.ntotal == 1998
    D, I = ix.search(X[:1], 2)
    assert 1995 in I[0].tolist()


def test_bad_arguments_fail_loudly(dev):
    from rag_arc_b200._native import RagArcError
    with pytest.raises(ValueError):
        NativeFlatIndex(64, "float32", "hamming")
    ix = NativeFlatIndex(64, "float32", "ip")
    ix.add(np.zeros((4, 64), np.float32))
    with pytest.raises(RagArcError):
        ix.remove([9])
    with pytest.raises(ValueError):
        ix.add(np.zeros((4, 32), np.float32))


@pytest.mark.parametrize("dtype,metric", [("float32", "ip"), ("bfloat16", "cosine"), ("float32", "l2"), ("float16", "l2")])
def test_single_process_sharded_index_equals_flat_index(dev, dtype, metric):
    """ragarc_sharded_*: three shards (all on device 0 here; tests/test_gpu_multi.py spreads them over
    two GPUs) must return exactly what one flat index returns - same scores bit for bit, same ids."""
    from rag_arc_b200.native_index import NativeShardedIndex
    X = synth.dense_corpus_np(20_001, 128, seed=5) * 2.5
    Q, _ = synth.dense_queries_np(X, 130, seed=6)
    flat = NativeFlatIndex(128, dtype, metric); flat.add(X)
    sh = NativeShardedIndex(128, dtype, metric, devices=(0, 0, 0)); sh.add(X)
    assert sh.ntotal == flat.ntotal == 20_001
    for k in (1, 10, 100):
        D, I = flat.search(Q, k); Ds, Is = sh.search(Q, k)
        assert np.array_equal(I, Is) and np.array_equal(D.view(np.uint32), Ds.view(np.uint32))
    # rows added later extend the last shard; a tiny first load leaves shards empty
    more = synth.dense_corpus_np(500, 128, seed=7) * 2.5
    flat.add(more); sh.add(more)
    D, I = flat.search(more[:9], 5); Ds, Is = sh.search(more[:9], 5)
    assert np.array_equal(I, Is) and np.array_equal(D, Ds) and (I[:, 0] >= 20_001).all()
    tiny = NativeShardedIndex(128, dtype, metric, devices=(0, 0, 0, 0)); tiny.add(X[:2])
    ft = NativeFlatIndex(128, dtype, metric); ft.add(X[:2])
    D, I = ft.search(Q[:3], 4); Ds, Is = tiny.search(Q[:3], 4)
    assert np.array_equal(I, Is) and (Is[:, 2:] == -1).all() and np.array_equal(D[:, :2], Ds[:, :2])
    tiny.add(X[2:40]); ft.add(X[2:40])
    D, I = ft.search(Q[:3], 4); Ds, Is = tiny.search(Q[:3], 4)
    assert np.array_equal(I, Is) and np.array_equal(D, Ds)
    for o in (flat, sh, tiny, ft):
        o.close()


@pytest.mark.parametrize("dtype", ["float32", "bfloat16", "float16"])
def test_l2_index_matches_indexflatl2_restatement(dev, dtype):
    """ragarc_index_* with RAGARC_METRIC_L2 against oracle.dense.IndexFlatL2 (squared distances
    ascending) on the values the index stores (storage-dtype rounding applied to both sides), incl.
    un-normalised rows of very different norms, remove, and k > ntotal padding."""
    import torch
    rng = np.random.default_rng(5)
    n, d, nq, k = 20_000, 100, 37, 10                       # d % 8 != 0 for fp32 (SIMT), augmented width pads for half types
    X = (rng.standard_normal((n, d)) * rng.uniform(0.2, 3.0, size=(n, 1))).astype(np.float32)
    Q = (X[rng.integers(0, n, nq)] + 0.1 * rng.standard_normal((nq, d))).astype(np.float32)
    tdt = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}[dtype]
    Xr = torch.from_numpy(X).to(tdt).float().numpy(); Qr = torch.from_numpy(Q).to(tdt).float().numpy()
    ref = odense.IndexFlatL2(d); ref.add(Xr)
    idx = NativeFlatIndex(d, dtype, "l2"); idx.add(X)
    D, I = idx.search(Q, k)
    Dr, Ir = ref.search(Qr, k)
    assert (np.diff(D, axis=1) >= 0).all()
    assert np.allclose(D, Dr, rtol=1e-5, atol=1e-4 * max(1.0, float(Dr.max())))
    assert (I == Ir).mean() > 0.99
    true = ((Qr[:, None, :].astype(np.float64) - Xr[I].astype(np.float64)) ** 2).sum(-1)
    assert np.allclose(D, true, rtol=1e-4, atol=1e-3)      # the returned rows really are at that distance
    idx.remove(I[:, 0].tolist())
    D2, I2 = idx.search(Q[:3], 3)
    assert (D2[:, 0] >= D[:3, 0] - 1e-6).all() and idx.ntotal == n - len(set(I[:, 0].tolist()))
    small = NativeFlatIndex(d, dtype, "l2"); small.add(X[:4])
    D3, I3 = small.search(Q[:2], 6)
    assert (I3[:, 4:] == -1).all() and np.isinf(D3[:, 4:]).all() and sorted(I3[0, :4].tolist()) == [0, 1, 2, 3]
