"""Scoring tail of the reference's Qwen3 reranker (/root/reference core/rerank/Reranker_Qwen3.py:
44-49): after the LM forward, the "yes"/"no" logits of the last position go through a two-way
log-softmax and the exp of the "yes" entry is the relevance score.  The LM forward itself is out
of scope; this keeps the tail on the device, one kernel for the batch."""
from __future__ import annotations

from typing import List

import torch

from ... import ops


def compute_scores_from_logits(logits: torch.Tensor, token_true_id: int, token_false_id: int) -> List[float]:
    """logits: ``[B, T, vocab]`` LM output (CUDA) -> list of B floats, as ``compute_logits`` returns."""
    last = logits[:, -1, :]
    return ops.yes_no_score(last, token_true_id, token_false_id).cpu().tolist()
