"""The dense oracle against an independent implementation that IS installed here: scikit-learn's brute-force
NearestNeighbors (the reference itself leans on sklearn's cosine similarity for its self-join,
encapsulation/database/graph_db/Base_Neo4j.py:538-590).  FAISS is absent from this image, so this does not pin
the oracle to the reference's own dependency - it shows that oracle/dense.py returns what a second exact
nearest-neighbour implementation returns on the same fp32 data: same neighbours in the same order on
tie-free inputs, same distances within fp32 rounding."""
import numpy as np
import pytest

from oracle import dense as odense

sk = pytest.importorskip("sklearn.neighbors")


def _data(n, d, nq, seed):
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, d)) * rng.uniform(0.5, 2.0, size=(n, 1))).astype(np.float32)
    Q = (X[rng.integers(0, n, nq)] + 0.2 * rng.standard_normal((nq, d))).astype(np.float32)
    return X, Q


@pytest.mark.parametrize("n,d,nq,k", [(2000, 64, 50, 10), (5000, 384, 20, 25)])
def test_cosine_search_equals_sklearn_brute_force(n, d, nq, k):
    X, Q = _data(n, d, nq, 1)
    Xn, Qn = X.copy(), Q.copy()
    odense.normalize_L2(Xn); odense.normalize_L2(Qn)             # the reference's cosine path: normalise, then IP
    D, I = odense.flat_ip_search(Xn, Qn, k)
    nn = sk.NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(X)
    dist, idx = nn.kneighbors(Q)                                  # cosine distance = 1 - cosine similarity, ascending
    assert np.array_equal(I, idx)
    assert np.allclose(1.0 - D, dist, atol=2e-6)


@pytest.mark.parametrize("n,d,nq,k", [(2000, 64, 50, 10), (5000, 100, 20, 25)])
def test_l2_search_equals_sklearn_brute_force(n, d, nq, k):
    X, Q = _data(n, d, nq, 2)
    ix = odense.IndexFlatL2(d); ix.add(X)
    D, I = ix.search(Q, k)                                        # squared distances, ascending (IndexFlatL2)
    nn = sk.NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean").fit(X.astype(np.float64))
    dist, idx = nn.kneighbors(Q.astype(np.float64))
    assert np.array_equal(I, idx)
    assert np.allclose(D, dist ** 2, rtol=1e-4, atol=1e-4)


def test_inner_product_search_equals_full_sort_in_float64():
    """Un-normalised inner product: sklearn has no such metric; the independent statement is a full argsort
    of the float64 score matrix."""
    X, Q = _data(3000, 96, 40, 3)
    D, I = odense.flat_ip_search(X, Q, 15)
    S = Q.astype(np.float64) @ X.astype(np.float64).T
    want = np.argsort(-S, axis=1, kind="stable")[:, :15]
    assert np.array_equal(I, want)
    assert np.allclose(D, np.take_along_axis(S, want, 1), rtol=1e-5, atol=1e-5)
