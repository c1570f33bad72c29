"""Seeded synthetic inputs of the shapes BASELINE.json names.

Dense corpora: rows ~ N(0,1) fp32 -> L2-normalised -> cast to the storage dtype; queries are
planted near random corpus rows (q = normalize(x[row] + 0.5*g), g ~ N(0, I/d)) so that every query
has a known nearest neighbour.  BM25 corpora: Zipf(1.07) token ids over a 50k vocabulary,
log-normal document lengths, queries of 8 tokens sampled from one document.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


def dense_corpus_np(n: int, d: int, seed: int = 1234) -> np.ndarray:
    """fp32 [n,d], rows L2-normalised (numpy; used for the CPU-sized configuration C1)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def dense_queries_np(x: np.ndarray, nq: int, seed: int = 4321) -> Tuple[np.ndarray, np.ndarray]:
    rng = np.random.default_rng(seed)
    n, d = x.shape
    rows = rng.integers(0, n, size=nq)
    g = rng.standard_normal((nq, d), dtype=np.float32) / np.sqrt(np.float32(d))
    q = x[rows].astype(np.float32) + np.float32(0.5) * g
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float32), rows


def dense_corpus_cuda(n: int, d: int, dtype: torch.dtype, device, seed: int = 1234,
                      chunk: int = 1 << 18, out: torch.Tensor = None) -> torch.Tensor:
    """[n,d] of ``dtype`` generated chunk by chunk on the device (never materialises fp32 whole)."""
    x = out if out is not None else torch.empty((n, d), dtype=dtype, device=device)
    gen = torch.Generator(device=device)
    for i, s in enumerate(range(0, n, chunk)):
        e = min(n, s + chunk)
        gen.manual_seed(seed + i)
        blk = torch.randn((e - s, d), generator=gen, device=device, dtype=torch.float32)
        blk = torch.nn.functional.normalize(blk, dim=1)
        x[s:e] = blk.to(dtype)
    return x


def dense_queries_cuda(x: torch.Tensor, nq: int, seed: int = 4321, n_rows: int = None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    n = x.shape[0] if n_rows is None else n_rows
    d = x.shape[1]
    gen = torch.Generator(device=x.device)
    gen.manual_seed(seed)
    rows = torch.randint(0, n, (nq,), generator=gen, device=x.device)
    g = torch.randn((nq, d), generator=gen, device=x.device, dtype=torch.float32) / (d ** 0.5)
    q = x[rows].float() + 0.5 * g
    q = torch.nn.functional.normalize(q, dim=1)
    return q.to(x.dtype), rows


def bm25_corpus_tokens(n_docs: int = 100_000, vocab: int = 50_000, seed: int = 777,
                       mean_len: float = 120.0, sigma: float = 0.4, zipf_s: float = 1.07
                       ) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (flat token ids int32, doc offsets int64[n_docs+1])."""
    rng = np.random.default_rng(seed)
    lens = np.maximum(8, np.rint(rng.lognormal(np.log(mean_len), sigma, size=n_docs))).astype(np.int64)
    offs = np.zeros(n_docs + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    p = ranks ** (-zipf_s)
    p /= p.sum()
    cdf = np.cumsum(p)
    u = rng.random(int(offs[-1]))
    toks = np.searchsorted(cdf, u).astype(np.int32)
    np.minimum(toks, vocab - 1, out=toks)
    return toks, offs


def bm25_queries_tokens(toks: np.ndarray, offs: np.ndarray, nq: int = 256, qlen: int = 8,
                        seed: int = 778) -> np.ndarray:
    """[nq, qlen] token ids, each query sampled (with replacement) from one random document."""
    rng = np.random.default_rng(seed)
    n_docs = len(offs) - 1
    docs = rng.integers(0, n_docs, size=nq)
    out = np.empty((nq, qlen), np.int32)
    for i, di in enumerate(docs):
        seg = toks[offs[di]:offs[di + 1]]
        out[i] = seg[rng.integers(0, len(seg), size=qlen)]
    return out


def tokens_to_texts(toks: np.ndarray, offs: np.ndarray) -> List[str]:
    """Space-joined ``t<i>`` tokens so the reference's ``str.split`` tokenizer applies."""
    names = np.char.add("t", np.arange(int(toks.max()) + 1).astype(str))
    return [" ".join(names[toks[offs[i]:offs[i + 1]]]) for i in range(len(offs) - 1)]
