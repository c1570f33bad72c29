"""TEST INFRASTRUCTURE - CPU restatements of two small tails next to the hot path.

* ``adjacent_cosine_distances``: /root/reference core/file_management/chunker/spliter.py:307-333
  (numpy fallback of ``cosine_similarity``: fp64 dot / outer(norms), nan/inf -> 0) applied to
  consecutive rows as ``calculate_cosine_distances`` (:354-372) does.  Pinned against the live
  reference function by ``oracle/gen_golden.py`` -> ``tests/golden/adjacent_cosine.npz``.
* ``yes_no_scores``: /root/reference core/rerank/Reranker_Qwen3.py:44-49 restated with the same
  torch calls on the CPU (the reference itself is torch code).
"""
from __future__ import annotations

import numpy as np


def adjacent_cosine_distances(X) -> np.ndarray:
    X = np.asarray(X, dtype=np.float64)
    out = np.empty(max(len(X) - 1, 0), np.float64)
    for i in range(len(X) - 1):
        a, b = X[i:i + 1], X[i + 1:i + 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            sim = np.dot(a, b.T) / np.outer(np.linalg.norm(a, axis=1), np.linalg.norm(b, axis=1))
        sim[np.isnan(sim) | np.isinf(sim)] = 0.0
        out[i] = 1 - sim[0][0]
    return out


def yes_no_scores(last_logits, true_id: int, false_id: int):
    import torch
    true_vector = last_logits[:, true_id]
    false_vector = last_logits[:, false_id]
    batch_scores = torch.stack([false_vector, true_vector], dim=1)
    batch_scores = torch.nn.functional.log_softmax(batch_scores, dim=1)
    return batch_scores[:, 1].exp().tolist()
