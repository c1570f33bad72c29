"""GPU: the two small tails next to the hot path - adjacent-row cosine
distances of the semantic chunker and the reranker's yes/no score - against the live-reference
golden vectors and the torch restatement of the reference's own torch code."""
import os

import numpy as np
import pytest
import torch

from oracle import chunk as ochunk
from rag_arc_b200 import ops
from rag_arc_b200.core.file_management.chunker.spliter import calculate_cosine_distances
from rag_arc_b200.core.rerank.scoring import compute_scores_from_logits

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_adjacent_cosine_matches_reference_golden(dev):
    z = np.load(os.path.join(GOLD, "adjacent_cosine.npz"))
    emb, want = z["emb"], z["distances"]
    for t in (torch.from_numpy(emb), torch.from_numpy(emb.astype(np.float64))):
        got = ops.adjacent_cosine_distance(t.to(dev)).cpu().numpy()
        # fp64 throughout; only the summation order of the dot products differs from numpy's
        assert np.allclose(got, want, rtol=0, atol=1e-13)
        assert got[6] == 1.0 and got[7] == 1.0          # zero row: nan similarity -> 0
    sentences = [{"sentence": str(i), "combined_sentence_embedding": e.tolist()} for i, e in enumerate(emb)]
    dist, sentences = calculate_cosine_distances(sentences, device=dev)
    assert np.allclose(dist, want, rtol=0, atol=1e-13)
    assert sentences[3]["distance_to_next"] == dist[3] and "distance_to_next" not in sentences[-1]
    assert ops.adjacent_cosine_distance(torch.zeros((1, 8), device=dev)).numel() == 0


def test_adjacent_cosine_large_random_against_oracle(dev):
    rng = np.random.default_rng(5)
    X = rng.standard_normal((3000, 1024)).astype(np.float32)
    got = ops.adjacent_cosine_distance(torch.from_numpy(X).to(dev)).cpu().numpy()
    want = ochunk.adjacent_cosine_distances(X)
    assert np.allclose(got, want, rtol=0, atol=1e-12)


@pytest.mark.parametrize("dtype,ulp", [(torch.float32, 2e-6), (torch.bfloat16, 2 ** -7), (torch.float16, 2 ** -10)])
def test_yes_no_score_matches_torch_restatement(dev, dtype, ulp):
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn((37, 5, 1000), generator=g) * 6).to(dtype)
    true_id, false_id = 812, 17
    want = np.asarray(ochunk.yes_no_scores(logits[:, -1, :], true_id, false_id), dtype=np.float64)
    got = np.asarray(compute_scores_from_logits(logits.to(dev), true_id, false_id), dtype=np.float64)
    assert got.shape == want.shape
    # both sides round the log-softmax and the exp to `dtype`; allow one unit in the last place
    assert np.all(np.abs(got - want) <= ulp * np.maximum(np.abs(want), 2.0 ** -14))
    assert np.all((got >= 0) & (got <= 1))
