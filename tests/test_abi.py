"""CPU: the C-ABI library builds, loads, and exports exactly what include/ragarc_b200.h declares;
entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ragarc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ragarc_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_declare_the_same_symbols():
    from rag_arc_b200 import _native
    assert _declared_symbols() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    from rag_arc_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert _native.lib.ragarc_abi_version() == 1


def test_signatures_use_plain_c_types_only():
    text = open(os.path.join(ROOT, "include", "ragarc_b200.h")).read()
    assert "torch" not in text and "at::" not in text and "std::" not in text
    assert 'extern "C"' in text


def test_workspace_queries_need_no_device():
    from rag_arc_b200 import _native as N
    ws = N.lib.ragarc_dense_topk_workspace_bytes(1_000_000, 768, N.BF16, 1024, 100)
    assert 50e6 < ws < 2e9
    assert N.lib.ragarc_bm25_workspace_bytes(100_000, 256) >= 100_000 * 8
    assert N.lib.ragarc_dense_topk_workspace_bytes(10, 8, N.BF16, 1, 5000) == 0     # k too large: unsupported


def test_invalid_arguments_return_error_codes_and_messages():
    from rag_arc_b200 import _native as N
    rc = N.lib.ragarc_rrf_fuse(None, 0, 1, 1, 60.0, 1, None, None, None, None)
    assert rc == 1 and b"rrf_fuse" in N.lib.ragarc_last_error()
    with pytest.raises(N.RagArcError):
        N.check(N.lib.ragarc_normalize_cast(None, None, -1, 0, 0, 1, None), "normalize_cast")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_ops_refuse_cpu_tensors_and_there_is_no_fallback():
    from rag_arc_b200 import _native as N
    from rag_arc_b200 import ops
    x = torch.zeros((4, 8)); q = torch.zeros((1, 8))
    with pytest.raises(N.RagArcError, match="CUDA"):
        ops.dense_topk(x, q, 2)
    with pytest.raises(N.RagArcError, match="CUDA"):
        ops.normalize_cast(x)
    # a raw call without a device must fail with the CUDA error code, not compute anything
    buf = (ctypes.c_float * 32)()
    rc = N.lib.ragarc_normalize_cast(ctypes.addressof(buf), ctypes.addressof(buf), 4, 8, N.F32, 1, None)
    assert rc == 2


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "rag_arc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_dense_plan_introspection_without_a_gpu():
    """ragarc_dense_topk_plan is pure host logic: it must answer without a device (148 SMs assumed)
    and obey its own invariants; unsupported shapes come back as error codes with a message."""
    from rag_arc_b200 import _native as N
    p = N.dense_plan(1_000_000, 768, N.BF16, 1024, 100)
    assert p["path"] == N.DENSE_TCGEN05 and p["rows_per_item"] == 256 and p["query_blocks"] == 4
    # thresholds come from published order statistics (>= k rows behind the minimum), not a seed pass
    assert p["slices"] * 100 <= 16384 and p["seed_rows"] == 0
    assert 0 < p["publishing_lists"] <= 32 and 0 < p["published_rank"] <= 8
    assert p["publishing_lists"] * p["published_rank"] >= 100
    assert 0 <= p["tail_slices"] < p["slices"] and 0 < p["cluster_tiles"] <= 3907
    p1 = N.dense_plan(1_000_000, 768, N.BF16, 1, 100)
    assert p1["rows_per_item"] == 128 and p1["pairs_per_cluster"] == 1 and p1["query_blocks"] == 1
    pf = N.dense_plan(10_000, 384, N.F32, 100, 10)
    assert pf["path"] == N.DENSE_SIMT and pf["rows_per_item"] == 64
    out = (ctypes.c_int * 16)()
    assert N.lib.ragarc_dense_topk_plan(1000, 64, N.BF16, 4, 5000, 0, out) == 4        # RAGARC_ERR_UNSUPPORTED
    assert b"2016" in N.lib.ragarc_last_error()
    assert N.lib.ragarc_dense_topk_plan(1000, 64, N.F32, 4, 5, N.DENSE_TCGEN05, out) == 4
