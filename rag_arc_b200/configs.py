"""Registry configs: ``AbstractConfig`` subclasses whose ``build()`` returns the B200 retrieval
plugins, so that ``Register().register(path, name, ConfigType)`` (framework/register.py) can stand
them up from a JSON file exactly as the reference's registry expects
(/root/reference framework/register.py:15-23, config.py:11-88, module.py:9-11).

Nested configs use pydantic discriminated unions on the literal ``type`` tag, the pattern the
reference's own tests exercise (framework/config_test.py:37-45, module_test.py:75-109).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Annotated, Any, Dict, List, Literal, Optional, Union

from pydantic import Field

from .framework.config import AbstractConfig
from .framework.module import AbstractModule


def _load_corpus(path: str):
    """JSON-lines (or a JSON list) of {"content": str, "metadata": {...}?, "id": str?}."""
    with open(path) as fh:
        head = fh.read(1)
        fh.seek(0)
        rows = json.load(fh) if head == "[" else [json.loads(line) for line in fh if line.strip()]
    texts = [r["content"] for r in rows]
    metas = [r.get("metadata", {}) for r in rows]
    ids = [r.get("id") for r in rows]
    return texts, metas, (ids if all(i is not None for i in ids) else None)


# ---- embeddings ------------------------------------------------------------------------------------
class HashEmbeddingsConfig(AbstractConfig):
    type: Literal["hash_embeddings"] = "hash_embeddings"
    dim: int = 384
    seed: int = 0

    def build(self):
        return EmbeddingsModule(config=self)


class PooledEncoderEmbeddingsConfig(AbstractConfig):
    type: Literal["b200_pooled_encoder"] = "b200_pooled_encoder"
    model_name: str
    pooling: Literal["mean", "cls", "last"] = "mean"
    normalize_embeddings: bool = True
    device: str = "cuda"
    batch_size: int = 64
    max_length: int = 512

    def build(self):
        return EmbeddingsModule(config=self)


class HuggingFaceEmbeddingsConfig(AbstractConfig):
    """The reference's ``HuggingFaceEmbeddings`` fields (core/file_management/embeddings/huggingface.py:67-83)
    as a registry config: builds the same-named class of this package."""
    type: Literal["huggingface_embeddings"] = "huggingface_embeddings"
    model_name: str = "sentence-transformers/all-mpnet-base-v2"
    cache_folder: Optional[str] = None
    model_kwargs: dict = Field(default_factory=dict)
    encode_kwargs: dict = Field(default_factory=dict)
    multi_process: bool = False
    show_progress_bar: bool = False

    def build(self):
        return EmbeddingsModule(config=self)


EmbeddingsConfig = Annotated[Union[HashEmbeddingsConfig, PooledEncoderEmbeddingsConfig, HuggingFaceEmbeddingsConfig],
                             Field(discriminator="type")]


@dataclass
class EmbeddingsModule(AbstractModule):
    impl: Any = field(default=None, init=False, repr=False)

    def __post_init__(self):
        from .encapsulation.embeddings.pooled import B200PooledEmbeddings, HashEmbeddings
        c = self.config
        if c.type == "hash_embeddings":
            self.impl = HashEmbeddings(dim=c.dim, seed=c.seed)
        elif c.type == "huggingface_embeddings":
            from .core.file_management.embeddings.huggingface import HuggingFaceEmbeddings
            self.impl = HuggingFaceEmbeddings(model_name=c.model_name, cache_folder=c.cache_folder,
                                              model_kwargs=c.model_kwargs, encode_kwargs=c.encode_kwargs,
                                              multi_process=c.multi_process, show_progress_bar=c.show_progress_bar)
        else:
            self.impl = B200PooledEmbeddings.from_pretrained(
                c.model_name, pooling=c.pooling, normalize_embeddings=c.normalize_embeddings, device=c.device,
                batch_size=c.batch_size, max_length=c.max_length)

    def embed_documents(self, texts):
        return self.impl.embed_documents(texts)

    def embed_query(self, text):
        return self.impl.embed_query(text)


# ---- vector store + dense retriever --------------------------------------------------------------------
class B200VectorStoreConfig(AbstractConfig):
    type: Literal["b200_vector_store"] = "b200_vector_store"
    embedding: EmbeddingsConfig
    metric: Literal["cosine", "ip"] = "cosine"
    dtype: Literal["float32", "float32x3", "float32_exact", "bfloat16", "float16"] = "float32"
    device: str = "cuda"
    corpus_path: Optional[str] = None       # JSON(L) corpus to index at build time
    index_path: Optional[str] = None        # folder written by save_local
    index_name: str = "index"

    def build(self):
        return VectorStoreModule(config=self)


@dataclass
class VectorStoreModule(AbstractModule):
    store: Any = field(default=None, init=False, repr=False)

    def __post_init__(self):
        from .encapsulation.database.vector_db.VectorStore_B200 import B200VectorStore
        c = self.config
        emb = c.embedding.build()
        if c.index_path:
            self.store = B200VectorStore.load_local(c.index_path, emb, c.index_name, device=c.device)
        else:
            self.store = B200VectorStore(embedding=emb, metric=c.metric, dtype=c.dtype, device=c.device)
            if c.corpus_path:
                texts, metas, ids = _load_corpus(c.corpus_path)
                self.store.add_texts(texts, metas, ids=ids)

    def __getattr__(self, name):          # delegate the VectorStore surface
        return getattr(self.__dict__["store"], name)


class DenseRetrieverConfig(AbstractConfig):
    type: Literal["b200_dense_retriever"] = "b200_dense_retriever"
    vectorstore: B200VectorStoreConfig
    search_type: Literal["similarity", "similarity_score_threshold", "mmr"] = "similarity"
    search_kwargs: Dict[str, Any] = Field(default_factory=dict)

    def build(self):
        return RetrieverModule(config=self)


class BM25RetrieverConfig(AbstractConfig):
    type: Literal["b200_bm25_retriever"] = "b200_bm25_retriever"
    corpus_path: str
    k: int = 5
    bm25_params: Dict[str, Any] = Field(default_factory=dict)
    device: str = "cuda"

    def build(self):
        return RetrieverModule(config=self)


class RRFusionConfig(AbstractConfig):
    type: Literal["rrf"] = "rrf"
    k: float = 60.0
    device: str = "cuda"

    def build(self):
        return FusionModule(config=self)


@dataclass
class FusionModule(AbstractModule):
    impl: Any = field(default=None, init=False, repr=False)

    def __post_init__(self):
        from .core.utils.Fusion import RRFusion
        self.impl = RRFusion(k=self.config.k, device=self.config.device)

    def fuse(self, results, top_k):
        return self.impl.fuse(results, top_k)


SubRetrieverConfig = Annotated[Union[DenseRetrieverConfig, BM25RetrieverConfig], Field(discriminator="type")]


class HybridRetrieverConfig(AbstractConfig):
    type: Literal["b200_hybrid_retriever"] = "b200_hybrid_retriever"
    retrievers: List[SubRetrieverConfig]
    fusion: RRFusionConfig = Field(default_factory=RRFusionConfig)
    top_k_per_retriever: int = 50

    def build(self):
        return RetrieverModule(config=self)


@dataclass
class RetrieverModule(AbstractModule):
    """Holds the built retriever; ``invoke`` / ``ainvoke`` / ``invoke_batch`` delegate to it."""
    retriever: Any = field(default=None, init=False, repr=False)

    def __post_init__(self):
        from .core.retrieval.bm25 import BM25Retriever
        from .core.retrieval.dense import VectorStoreRetriever
        from .core.retrieval.mutipath import MultiPathRetriever
        c = self.config
        if c.type == "b200_dense_retriever":
            store = c.vectorstore.build().store
            self.retriever = VectorStoreRetriever(vectorstore=store, search_type=c.search_type,
                                                  search_kwargs=dict(c.search_kwargs))
        elif c.type == "b200_bm25_retriever":
            texts, metas, ids = _load_corpus(c.corpus_path)
            self.retriever = BM25Retriever.from_texts(texts, metas, ids, bm25_params=dict(c.bm25_params), k=c.k,
                                                      device=c.device)
        else:
            subs = [sub.build().retriever for sub in c.retrievers]
            self.retriever = MultiPathRetriever(subs, fusion_method=c.fusion.build().impl,
                                                top_k_per_retriever=c.top_k_per_retriever)

    def invoke(self, query: str, **kwargs):
        return self.retriever.invoke(query, **kwargs)

    async def ainvoke(self, query: str, **kwargs):
        return await self.retriever.ainvoke(query, **kwargs)

    def invoke_batch(self, queries, **kwargs):
        return self.retriever.invoke_batch(queries, **kwargs)
