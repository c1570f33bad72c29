"""``HuggingFaceEmbeddings``: same name, constructor and behaviour as the reference class
(/root/reference core/file_management/embeddings/huggingface.py:9-145), without the
``sentence_transformers`` dependency on the hot path.

The reference builds ``sentence_transformers.SentenceTransformer(model_name, cache_folder=...,
**model_kwargs)`` (:95-98) and calls ``.encode(texts, show_progress_bar=..., **encode_kwargs)``
(:122-126), which tokenises, runs the transformer, pools as the model's ``1_Pooling/config.json``
says, normalises when a ``2_Normalize`` module is present or ``normalize_embeddings`` is passed, and
returns fp32 rows that become Python lists (:134).  Here the transformer forward stays the
HuggingFace ``AutoModel`` (out of scope of the B200 path), and everything after it - masked mean /
CLS / last-token pooling and the L2 normalisation - is the fused ``ragarc_pool_normalize`` kernel
(``B200PooledEmbeddings``).  A JSON / Python config written for the reference class constructs this
one unchanged: the same six fields, unknown fields rejected (the reference is a pydantic model with
``extra="forbid"``, :100-103).

Honoured ``model_kwargs``: ``device``, ``prompts``, ``default_prompt_name``, ``torch_dtype`` /
``model_kwargs.torch_dtype``, ``trust_remote_code``, ``local_files_only``; honoured ``encode_kwargs``:
``prompt_name``, ``prompt``, ``batch_size``, ``normalize_embeddings``.  ``multi_process`` is accepted
and answered by the single device (one process per GPU is this package's model; a pool of CPU
workers has nothing to add).  ``embed_documents_tensor`` additionally keeps the result on the
device for ``B200VectorStore.add_embeddings`` / ``search_batch``.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .base import Embeddings

DEFAULT_MODEL_NAME = "sentence-transformers/all-mpnet-base-v2"
_FIELDS = ("model_name", "cache_folder", "model_kwargs", "encode_kwargs", "multi_process", "show_progress_bar")


def _pooling_of(model_dir: str) -> tuple:
    """(pooling mode, has a Normalize module) as sentence-transformers reads them from a model
    directory: modules.json lists the modules, ``<path>/config.json`` of the Pooling module holds
    ``pooling_mode_*`` flags.  A plain transformers checkpoint (no modules.json) is mean-pooled,
    which is what SentenceTransformer falls back to."""
    mode, has_norm = "mean", False
    mj = os.path.join(model_dir, "modules.json")
    pool_dir = os.path.join(model_dir, "1_Pooling")
    if os.path.isfile(mj):
        with open(mj) as f:
            for mod in json.load(f):
                typ = str(mod.get("type", ""))
                if typ.endswith("Pooling"):
                    pool_dir = os.path.join(model_dir, mod.get("path", "1_Pooling"))
                if typ.endswith("Normalize"):
                    has_norm = True
    cfg = os.path.join(pool_dir, "config.json")
    if os.path.isfile(cfg):
        with open(cfg) as f:
            c = json.load(f)
        if c.get("pooling_mode_cls_token"):
            mode = "cls"
        elif c.get("pooling_mode_lasttoken"):
            mode = "last"
        elif c.get("pooling_mode_mean_tokens", True):
            mode = "mean"
        else:
            raise ValueError(f"{cfg}: only mean / cls / last-token pooling are offered")
    return mode, has_norm


class HuggingFaceEmbeddings(Embeddings):
    def __init__(self, **kwargs: Any):
        unknown = set(kwargs) - set(_FIELDS)
        if unknown:
            raise ValueError(f"extra fields not permitted: {sorted(unknown)}")      # pydantic extra="forbid"
        self.model_name: str = kwargs.get("model_name", DEFAULT_MODEL_NAME)
        self.cache_folder: Optional[str] = kwargs.get("cache_folder")
        self.model_kwargs: Dict[str, Any] = dict(kwargs.get("model_kwargs") or {})
        self.encode_kwargs: Dict[str, Any] = dict(kwargs.get("encode_kwargs") or {})
        self.multi_process: bool = bool(kwargs.get("multi_process", False))
        self.show_progress_bar: bool = bool(kwargs.get("show_progress_bar", False))
        from ....encapsulation.embeddings.pooled import B200PooledEmbeddings
        mk = dict(self.model_kwargs)
        device = mk.pop("device", "cuda")
        self.prompts: Dict[str, str] = dict(mk.pop("prompts", None) or {})
        self.default_prompt_name: Optional[str] = mk.pop("default_prompt_name", None)
        if self.default_prompt_name is not None and self.default_prompt_name not in self.prompts:
            raise ValueError(f"default_prompt_name {self.default_prompt_name!r} is not a key of prompts")
        inner = dict(mk.pop("model_kwargs", None) or {})
        dtype = mk.pop("torch_dtype", inner.pop("torch_dtype", torch.float16))
        if isinstance(dtype, str):
            dtype = getattr(torch, dtype.replace("torch.", ""))
        path = self.model_name
        if self.cache_folder and not os.path.isdir(path):
            cand = os.path.join(self.cache_folder, self.model_name.replace("/", "_"))
            path = cand if os.path.isdir(cand) else path
        mode, has_norm = _pooling_of(path) if os.path.isdir(path) else ("mean", False)
        self.pooling, self._always_normalize = mode, has_norm
        self._client = B200PooledEmbeddings.from_pretrained(
            path, pooling=mode, normalize_embeddings=True, device=device, dtype=dtype,
            batch_size=int(self.encode_kwargs.get("batch_size", 32)), cache_dir=self.cache_folder,
            trust_remote_code=bool(mk.pop("trust_remote_code", False)),
            local_files_only=bool(mk.pop("local_files_only", False)))

    # ---- the reference's two methods (lists of floats) ------------------------------------------
    def _prepare(self, texts: List[str]) -> List[str]:
        texts = [t.replace("\n", " ") for t in texts]                      # huggingface.py:116
        prompt = self.encode_kwargs.get("prompt")
        if prompt is None:
            name = self.encode_kwargs.get("prompt_name", self.default_prompt_name)
            if name is not None:
                if name not in self.prompts:
                    raise ValueError(f"prompt name {name!r} not found in the configured prompts {sorted(self.prompts)}")
                prompt = self.prompts[name]
        return [prompt + t for t in texts] if prompt else texts

    def embed_documents_tensor(self, texts: List[str]) -> torch.Tensor:
        """fp32 ``[n, H]`` on the device (no ``.tolist()`` round trip)."""
        self._client.normalize_embeddings = bool(self._always_normalize or self.encode_kwargs.get("normalize_embeddings", False))
        self._client.batch_size = int(self.encode_kwargs.get("batch_size", 32))
        return self._client.embed_documents_tensor(self._prepare(list(texts)))

    def embed_documents_array(self, texts: List[str]) -> np.ndarray:
        return self.embed_documents_tensor(texts).float().cpu().numpy()

    def embed_documents(self, texts: List[str]) -> List[List[float]]:
        return self.embed_documents_tensor(texts).cpu().tolist()            # :134

    def embed_query(self, text: str) -> List[float]:
        return self.embed_documents([text])[0]                              # :145
