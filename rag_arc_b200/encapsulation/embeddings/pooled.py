"""Embedding plugins whose pool + L2-normalise tail runs on the GPU.

The reference's ``HuggingFaceEmbeddings`` (/root/reference
core/file_management/embeddings/huggingface.py:85-145) hands tokenisation, the transformer,
pooling and normalisation to ``sentence_transformers`` and converts the result to Python lists
(:134).  Here the encoder forward stays a plain callable (out of scope), and the part of that call
that is arithmetic on the hot path - masked mean / CLS / last-token pooling followed by L2
normalisation - is the fused ``ragarc_pool_normalize`` kernel.  ``embed_documents_tensor`` keeps
the result on the device so it can feed ``B200VectorStore.add_embeddings`` / ``search_batch``
without the ``.tolist()`` round trip; ``embed_documents`` / ``embed_query`` keep the reference's
list-of-floats contract.
"""
from __future__ import annotations

import zlib
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from ...core.file_management.embeddings.base import Embeddings
from ... import ops


class B200PooledEmbeddings(Embeddings):
    """``encoder(texts) -> (hidden [B,T,H] CUDA tensor, attention_mask [B,T])``."""

    def __init__(self, encoder: Callable[[List[str]], Tuple[torch.Tensor, torch.Tensor]],
                 pooling: str = "mean", normalize_embeddings: bool = True, batch_size: int = 64, **kwargs):
        super().__init__(**kwargs)
        if pooling not in ("mean", "cls", "last"):
            raise ValueError("pooling must be 'mean', 'cls' or 'last'")
        self.encoder = encoder
        self.pooling = pooling
        self.normalize_embeddings = normalize_embeddings
        self.batch_size = batch_size

    def embed_documents_tensor(self, texts: List[str]) -> torch.Tensor:
        texts = [t.replace("\n", " ") for t in texts]           # huggingface.py:116
        outs = []
        for s in range(0, len(texts), self.batch_size):
            hidden, mask = self.encoder(texts[s:s + self.batch_size])
            outs.append(ops.pool_normalize(hidden.contiguous(), mask, self.pooling, self.normalize_embeddings))
        return torch.cat(outs, dim=0) if outs else torch.empty((0, 0))

    def embed_documents_array(self, texts: List[str]) -> np.ndarray:
        return self.embed_documents_tensor(list(texts)).float().cpu().numpy()

    def embed_documents(self, texts: List[str]) -> List[List[float]]:
        return self.embed_documents_tensor(texts).cpu().tolist()

    def embed_query(self, text: str) -> List[float]:
        return self.embed_documents([text])[0]

    @classmethod
    def from_pretrained(cls, model_name: str, pooling: str = "mean", normalize_embeddings: bool = True,
                        device="cuda", dtype=torch.float16, max_length: int = 512, cache_dir=None,
                        trust_remote_code: bool = False, local_files_only: bool = False, **kwargs):
        """Wrap a HuggingFace ``AutoModel`` (weights must be available locally)."""
        from transformers import AutoModel, AutoTokenizer
        hub = dict(cache_dir=cache_dir, trust_remote_code=trust_remote_code, local_files_only=local_files_only)
        tok = AutoTokenizer.from_pretrained(model_name, **hub)
        model = AutoModel.from_pretrained(model_name, dtype=dtype, **hub).to(device).eval()

        @torch.no_grad()
        def encoder(texts):
            enc = tok(texts, padding=True, truncation=True, max_length=max_length, return_tensors="pt").to(device)
            return model(**enc).last_hidden_state, enc["attention_mask"]

        self = cls(encoder, pooling=pooling, normalize_embeddings=normalize_embeddings, **kwargs)
        self.tokenizer, self.model = tok, model
        return self


class TableEmbeddings(Embeddings):
    """Pre-computed embeddings looked up by text (used when vectors come from an offline job)."""

    def __init__(self, table, **kwargs):
        super().__init__(**kwargs)
        self.table = table

    def embed_documents(self, texts: List[str]) -> List[List[float]]:
        return [np.asarray(self.table[t], dtype=np.float32).tolist() for t in texts]

    def embed_documents_array(self, texts: List[str]) -> np.ndarray:
        """Same vectors as one float32 ``[n,d]`` array (skips the list-of-floats round trip the
        reference interface imposes; batched retrievers use it when present): one fancy-index gather
        out of a matrix built on first use."""
        if getattr(self, "_matrix", None) is None or len(self._row_of) != len(self.table):
            keys = list(self.table)
            self._row_of = {t: i for i, t in enumerate(keys)}
            self._matrix = np.stack([np.asarray(self.table[t], dtype=np.float32) for t in keys]) if keys else np.zeros((0, 0), np.float32)
        return self._matrix[np.fromiter((self._row_of[t] for t in texts), np.int64, len(texts))]

    def embed_query(self, text: str) -> List[float]:
        return np.asarray(self.table[text], dtype=np.float32).tolist()


class HashEmbeddings(Embeddings):
    """Deterministic bag-of-tokens embedding: every token gets a fixed pseudo-random direction
    (seeded by its CRC32), a text is the sum of its tokens' directions.  No model weights needed;
    texts that share tokens are close.  For demos, smoke tests and registry examples."""

    def __init__(self, dim: int = 384, seed: int = 0, **kwargs):
        super().__init__(**kwargs)
        self.dim, self.seed = dim, seed
        self._cache = {}

    def _tok(self, tok: str) -> np.ndarray:
        v = self._cache.get(tok)
        if v is None:
            rng = np.random.default_rng((zlib.crc32(tok.encode("utf-8")) << 8) ^ self.seed)
            v = rng.standard_normal(self.dim).astype(np.float32)
            self._cache[tok] = v
        return v

    def _embed(self, text: str) -> np.ndarray:
        toks = text.split()
        if not toks:
            return np.zeros(self.dim, np.float32)
        return np.sum([self._tok(t) for t in toks], axis=0, dtype=np.float32)

    def embed_documents(self, texts: List[str]) -> List[List[float]]:
        return [self._embed(t).tolist() for t in texts]

    def embed_query(self, text: str) -> List[float]:
        return self._embed(text).tolist()
