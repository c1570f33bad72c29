"""The one numeric step of the reference's ``SemanticChunker`` that is worth a kernel: the cosine
distance between the embeddings of consecutive sentence groups
(/root/reference core/file_management/chunker/spliter.py:354-372, which calls the row-wise
``cosine_similarity`` of :307-333 once per pair from Python).  Same name, same arguments, same
side effect on the sentence dicts; the splitter logic around it stays host code in the reference.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .... import ops


def calculate_cosine_distances(sentences: List[dict], device="cuda") -> Tuple[List[float], List[dict]]:
    """``sentences[i]["combined_sentence_embedding"]`` -> distances between neighbours (fp64, like the
    reference's numpy path); also stored as ``sentences[i]["distance_to_next"]``."""
    if len(sentences) < 2:
        return [], sentences
    emb = np.asarray([s["combined_sentence_embedding"] for s in sentences], dtype=np.float64)
    dist = ops.adjacent_cosine_distance(torch.from_numpy(emb).to(device)).cpu().tolist()
    for i, dv in enumerate(dist):
        sentences[i]["distance_to_next"] = dv
    return dist, sentences
