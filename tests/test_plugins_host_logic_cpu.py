"""The plugin-level golden tests of tests/test_gpu_plugins.py run here a second time on the CPU with
the kernels replaced by oracle-backed stand-ins (tests/fake_ops.py): what is checked is the HOST
logic of the plugin classes - docstore bookkeeping, relevance-score maps and thresholds, the MMR and
hybrid plumbing, persistence, the registry - against the vectors the reference's own classes
produced.  (The same bodies run against the real library under ``-m gpu``.)"""
import pytest
import torch

import fake_ops
import test_gpu_plugins as T

CPU = torch.device("cpu")


@pytest.fixture(autouse=True)
def _fake_kernels():
    with fake_ops.patched():
        yield


@pytest.mark.parametrize("metric", ["cosine", "ip"])
def test_vector_store_host_logic_matches_reference_golden(metric):
    T.test_vector_store_matches_reference_faiss_store_golden(CPU, metric)


def test_vector_store_bookkeeping_delete_persist(tmp_path):
    T.test_vector_store_bookkeeping_delete_persist_batch(CPU, tmp_path)


def test_bm25_retriever_host_logic_matches_reference_golden():
    T.test_bm25_retriever_matches_reference_golden(CPU)


def test_bm25_persistence(tmp_path):
    T.test_bm25_persistence_roundtrip(CPU, tmp_path)


def test_rrfusion_plugin_semantics():
    T.test_rrfusion_plugin_matches_reference_semantics(CPU)


def test_hybrid_retriever_equals_reference_on_tie_free_golden():
    T.test_hybrid_retriever_equals_reference_multipath_on_tie_free_golden(CPU)


def test_hybrid_returns_the_reference_document_objects():
    T.test_hybrid_returns_the_documents_of_the_retriever_the_reference_returns_them_from(CPU)


def test_hybrid_retriever_with_ties_and_duplicates():
    T.test_hybrid_retriever_with_ties_and_duplicate_content_is_consistent_with_fusion_oracle(CPU)


def test_hybrid_batch_follows_an_update_that_keeps_the_corpus_size():
    T.test_hybrid_batch_follows_an_update_that_keeps_the_corpus_size(CPU)


def test_registry_builds_hybrid_retriever(tmp_path):
    """Same JSON as the GPU test, with every ``device`` field pointing at the CPU stand-ins."""
    import json
    from rag_arc_b200.configs import HybridRetrieverConfig
    from rag_arc_b200.framework import Register
    corpus = tmp_path / "corpus.jsonl"
    rows = [{"content": f"doc about topic{i % 5} number{i}", "id": f"id{i}"} for i in range(60)]
    corpus.write_text("\n".join(json.dumps(r) for r in rows))
    cfg = {"type": "b200_hybrid_retriever", "top_k_per_retriever": 20, "fusion": {"type": "rrf", "device": "cpu"},
           "retrievers": [
               {"type": "b200_dense_retriever",
                "vectorstore": {"type": "b200_vector_store", "embedding": {"type": "hash_embeddings", "dim": 64},
                                "dtype": "bfloat16", "corpus_path": str(corpus), "device": "cpu"}},
               {"type": "b200_bm25_retriever", "corpus_path": str(corpus), "device": "cpu"}]}
    path = tmp_path / "hybrid.json"
    path.write_text(json.dumps(cfg))
    reg = Register()
    reg.register(str(path), "hybrid", HybridRetrieverConfig)
    app = reg.get_object("hybrid")
    docs = app.invoke("topic3 number13", top_k=5)
    assert docs[0].id == "id13" and len(docs) == 5
    assert [d.id for d in app.invoke_batch(["topic3 number13"], top_k=5)[0]] == [d.id for d in docs]


def test_pooled_embeddings_plugin():
    T.test_pooled_embeddings_plugin(CPU)


def test_load_local_imports_reference_folder():
    T.test_load_local_imports_a_folder_saved_by_the_reference(CPU)


def test_load_local_bitwise_rows_and_folder_validation(tmp_path):
    T.test_load_local_keeps_reference_rows_bitwise_and_validates_the_folder(CPU, tmp_path)


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_vector_store_l2_metric_host_logic(dtype, tmp_path):
    T.test_vector_store_l2_metric_matches_reference_golden(CPU, dtype, tmp_path)


def test_huggingface_embeddings_host_logic(tmp_path):
    T.test_huggingface_embeddings_same_constructor_as_the_reference(CPU, tmp_path)
