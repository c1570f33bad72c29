"""Oracle (test infrastructure): pooling + L2 normalisation of encoder outputs.

The reference delegates to ``SentenceTransformer.encode(texts, **encode_kwargs)``
(/root/reference ``core/file_management/embeddings/huggingface.py:122-134``); which pooling runs
is set by the model (mean for all-mpnet-base-v2 / MiniLM, CLS for bge, last-token for Qwen3
embedding).  sentence-transformers is third-party, un-vendored and un-pinned: **parity
unpinned**; this restates its published ``Pooling`` / ``Normalize`` modules:

    mean : sum_t(m_t * x_t) / clamp(sum_t m_t, min=1e-9)
    cls  : x_0
    last : x_{(sum_t m_t) - 1}         (right-padded batches; left-padded -> x_{T-1})
    normalize : x / max(||x||_2, 1e-12)

Output fp32 (``huggingface.py:134`` converts to Python floats).
"""
from __future__ import annotations

import numpy as np

__all__ = ["pool_normalize"]


def pool_normalize(x: np.ndarray, mask: np.ndarray, mode: str = "mean", normalize: bool = True
                   ) -> np.ndarray:
    """x: [B,T,H] (any float dtype, values taken exactly), mask: [B,T] of 0/1."""
    x32 = x.astype(np.float32)
    m = mask.astype(np.float32)
    if mode == "mean":
        s = (x32 * m[:, :, None]).sum(axis=1, dtype=np.float32)
        cnt = np.maximum(m.sum(axis=1, dtype=np.float32), np.float32(1e-9))
        out = s / cnt[:, None]
    elif mode == "cls":
        out = x32[:, 0, :].copy()
    elif mode == "last":
        B, T, _ = x32.shape
        left_padded = bool(mask[:, -1].sum() == B)
        if left_padded:
            out = x32[:, -1, :].copy()
        else:
            idx = mask.astype(np.int64).sum(axis=1) - 1
            out = x32[np.arange(B), idx, :].copy()
    else:
        raise ValueError(mode)
    if normalize:
        nrm = np.sqrt((out.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))
        out = out / np.maximum(nrm, np.float32(1e-12))[:, None]
    return out.astype(np.float32)
