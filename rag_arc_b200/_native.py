"""ctypes binding of ``libragarc_b200.so`` (the C ABI declared in ``include/ragarc_b200.h``).

There is deliberately no fallback: if the shared library is missing, importing this module fails
with instructions to build it, and every compute call raises if CUDA is unavailable.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from ctypes import c_double, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# RAGARC_LIB: load another build of the same library (e.g. one compiled with -DRAGARC_TC_STATS_BUILD)
LIB_PATH = os.environ.get("RAGARC_LIB") or os.path.join(_HERE, "libragarc_b200.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["common.cu", "merge.cu", "dense_simt.cu", "dense_tc.cu", "bm25.cu", "misc.cu", "index.cu", "vocab.cu", "comm.cu"]

F32, BF16, F16, F64 = 0, 1, 2, 3
METRIC_IP, METRIC_COSINE, METRIC_L2 = 0, 1, 2
POOL_MEAN, POOL_CLS, POOL_LAST = 0, 1, 2
DENSE_AUTO, DENSE_SIMT, DENSE_TCGEN05 = 0, 1, 2

# every symbol include/ragarc_b200.h declares (tests/test_abi.py checks the two stay in sync)
EXPORTS = [
    "ragarc_abi_version", "ragarc_last_error", "ragarc_launch_count", "ragarc_profile_enable",
    "ragarc_profile_read", "ragarc_normalize_cast",
    "ragarc_dense_topk_workspace_bytes", "ragarc_dense_topk_plan", "ragarc_dense_topk", "ragarc_dense_topk_keys",
    "ragarc_dense_topk_keys_push", "ragarc_merge_topk_inbox", "ragarc_dense_topk_ex",
    "ragarc_l2_aug_dim", "ragarc_l2_augment", "ragarc_l2_distances",
    "ragarc_normalize_split3", "ragarc_dense_topk_x3_workspace_bytes", "ragarc_dense_topk_x3",
    "ragarc_merge_topk_keys", "ragarc_merge_topk_keys_p2p", "ragarc_bm25_workspace_bytes", "ragarc_bm25_scores",
    "ragarc_bm25_topk", "ragarc_bm25_merge_topk", "ragarc_vocab_create", "ragarc_vocab_free", "ragarc_vocab_size",
    "ragarc_vocab_encode_split", "ragarc_vocab_encode_split0", "ragarc_rrf_fuse", "ragarc_rrf_fuse_rows", "ragarc_pool_normalize", "ragarc_mmr_select",
    "ragarc_adjacent_cosine_distance", "ragarc_yes_no_score",
    "ragarc_host_alloc", "ragarc_host_free", "ragarc_index_create", "ragarc_index_free", "ragarc_index_reserve", "ragarc_index_add", "ragarc_index_search",
    "ragarc_index_remove", "ragarc_index_ntotal", "ragarc_index_dim", "ragarc_index_rows",
    "ragarc_comm_nccl_version", "ragarc_comm_unique_id", "ragarc_comm_init_rank", "ragarc_comm_init_all",
    "ragarc_comm_free", "ragarc_comm_rank", "ragarc_comm_nranks", "ragarc_sharded_topk_workspace_bytes",
    "ragarc_sharded_topk",
    "ragarc_sharded_create", "ragarc_sharded_free", "ragarc_sharded_add", "ragarc_sharded_search",
    "ragarc_sharded_ntotal",
]


def nvcc_command(out: str = LIB_PATH):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177,550,128", "-ldl",
            *[os.path.join(CSRC, s) for s in SOURCES], "-o", out]


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into the in-tree shared library."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, "common.cuh"),
                                                       os.path.join(_HERE, "..", "include", "ragarc_b200.h")]
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in srcs)
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. rag_arc_b200 has no CPU/PyTorch fallback: build the CUDA "
            "library first with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or rag_arc_b200._native.build_library()).")
    lib = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    sig = {
        "ragarc_abi_version": (c_int, []),
        "ragarc_last_error": (ctypes.c_char_p, []),
        "ragarc_launch_count": (c_uint64, []),
        "ragarc_profile_enable": (c_int, [c_int]),
        "ragarc_profile_read": (c_int, [ctypes.POINTER(c_double), ctypes.POINTER(c_double),
                                        ctypes.POINTER(c_double), ctypes.POINTER(c_int)]),
        "ragarc_normalize_cast": (c_int, [P, P, c_int64, c_int, c_int, c_int, P]),
        "ragarc_dense_topk_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int, c_int]),
        "ragarc_dense_topk_plan": (c_int, [c_int64, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int)]),
        "ragarc_dense_topk": (c_int, [P, c_int64, c_int, c_int, P, c_int, c_int, P, P, P, c_size_t,
                                      c_int, ctypes.POINTER(c_int), P]),
        "ragarc_dense_topk_keys": (c_int, [P, c_int64, c_int, c_int, P, c_int, c_int, c_uint64, P, P,
                                           c_size_t, c_int, ctypes.POINTER(c_int), P]),
        "ragarc_dense_topk_keys_push": (c_int, [P, c_int64, c_int, c_int, P, c_int, c_int, c_uint64, P, c_int,
                                                c_int, c_int, c_int, P, c_size_t, c_int, ctypes.POINTER(c_int), P]),
        "ragarc_dense_topk_ex": (c_int, [P, c_int64, c_int, c_int, P, c_int, c_int, P, P, c_size_t, c_int,
                                         ctypes.POINTER(c_int), P]),
        "ragarc_merge_topk_inbox": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, P, c_double, P, P]),
        "ragarc_l2_aug_dim": (c_int, [c_int, c_int]),
        "ragarc_l2_augment": (c_int, [P, P, c_int64, c_int, c_int, c_int, c_int, P, P]),
        "ragarc_l2_distances": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
        "ragarc_normalize_split3": (c_int, [P, P, c_int64, c_int, c_int, P]),
        "ragarc_dense_topk_x3_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
        "ragarc_dense_topk_x3": (c_int, [P, c_int64, c_int, P, c_int, c_int, P, P, P, c_size_t, P]),
        "ragarc_merge_topk_keys": (c_int, [P, c_int, c_int, c_int, c_int, P, P, P]),
        "ragarc_merge_topk_keys_p2p": (c_int, [P, c_int, c_int, c_int, c_int, P, P, P]),
        "ragarc_bm25_workspace_bytes": (c_size_t, [c_int64, c_int]),
        "ragarc_bm25_scores": (c_int, [P, P, P, P, P, P, c_double, P, P, c_int, c_int, c_int64, P, P]),
        "ragarc_adjacent_cosine_distance": (c_int, [P, c_int, c_int64, c_int, P, P]),
        "ragarc_yes_no_score": (c_int, [P, c_int, c_int, c_int64, c_int, c_int, c_int, P, P]),
        "ragarc_host_alloc": (c_int, [c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
        "ragarc_host_free": (c_int, [P]),
        "ragarc_index_create": (c_int, [c_int, c_int, c_int, ctypes.POINTER(ctypes.c_void_p)]),
        "ragarc_index_free": (c_int, [P]),
        "ragarc_index_reserve": (c_int, [P, c_int64, P]),
        "ragarc_index_add": (c_int, [P, P, c_int64, c_int, P]),
        "ragarc_index_search": (c_int, [P, P, c_int, c_int, P, P, c_int, P]),
        "ragarc_index_remove": (c_int, [P, P, c_int64, P]),
        "ragarc_index_ntotal": (c_int64, [P]),
        "ragarc_index_dim": (c_int, [P]),
        "ragarc_index_rows": (ctypes.c_void_p, [P]),
        "ragarc_comm_nccl_version": (c_int, []),
        "ragarc_comm_unique_id": (c_int, [P]),
        "ragarc_comm_init_rank": (c_int, [P, c_int, c_int, ctypes.POINTER(ctypes.c_void_p)]),
        "ragarc_comm_init_all": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_void_p)]),
        "ragarc_comm_free": (c_int, [P]),
        "ragarc_comm_rank": (c_int, [P]),
        "ragarc_comm_nranks": (c_int, [P]),
        "ragarc_sharded_topk_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int, c_int, c_int]),
        "ragarc_sharded_topk": (c_int, [P, P, c_int64, c_int, c_int, P, c_int, c_int, c_uint64, P, P, P, c_size_t, P]),
        "ragarc_sharded_create": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_void_p)]),
        "ragarc_sharded_free": (c_int, [P]),
        "ragarc_sharded_add": (c_int, [P, P, c_int64]),
        "ragarc_sharded_search": (c_int, [P, P, c_int, c_int, P, P]),
        "ragarc_sharded_ntotal": (c_int64, [P]),
        "ragarc_bm25_merge_topk": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P]),
        "ragarc_vocab_create": (c_int, [P, P, c_int64, ctypes.POINTER(ctypes.c_void_p)]),
        "ragarc_vocab_free": (c_int, [P]),
        "ragarc_vocab_size": (c_int64, [P]),
        "ragarc_vocab_encode_split": (c_int, [P, P, P, c_int, c_int, P, P, ctypes.POINTER(c_int)]),
        "ragarc_vocab_encode_split0": (c_int, [P, P, c_int64, c_int, c_int, P, P, ctypes.POINTER(c_int)]),
        "ragarc_bm25_topk": (c_int, [P, P, P, P, P, P, c_double, P, P, c_int, c_int, c_int64, c_int, P, P,
                                     P, c_size_t, P]),
        "ragarc_rrf_fuse": (c_int, [P, c_int, c_int, c_int, c_double, c_int, P, P, P, P]),
        "ragarc_rrf_fuse_rows": (c_int, [P, P, P, c_int, c_int, c_int, c_double, c_int, P, P, P, P, P, P]),
        "ragarc_pool_normalize": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, P, P]),
        "ragarc_mmr_select": (c_int, [P, c_int64, c_int, c_int, P, c_int, P, c_int, c_int, c_double, P, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ragarc_abi_version() != 1:
        raise ImportError("libragarc_b200.so ABI version mismatch; rebuild it")
    return lib


lib = _load()


PHASE_BOTH, PHASE_SCORE, PHASE_SELECT = 0, 1, 2


class DenseOpts(ctypes.Structure):
    """``ragarc_dense_opts_t`` (include/ragarc_b200.h)."""
    _fields_ = [("phase", c_int), ("id_base", c_uint64), ("out_keys", c_void_p), ("out_scores", c_void_p),
                ("out_ids", c_void_p), ("inboxes", c_void_p), ("n_ranks", c_int), ("rank", c_int),
                ("nq_per_rank", c_int), ("signal", c_int), ("workspace_clean", c_int)]


class RagArcError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.ragarc_last_error().decode("utf-8", "replace")
        raise RagArcError(f"{what or 'ragarc'} failed (code {rc}): {msg}")


def profile_enable(on: bool) -> None:
    check(lib.ragarc_profile_enable(int(bool(on))), "profile_enable")


def profile_read():
    """-> (seed ms, scoring kernel ms, merge kernel ms - each summed - and the number of searches)."""
    s, a, b, n = c_double(0), c_double(0), c_double(0), c_int(0)
    check(lib.ragarc_profile_read(ctypes.byref(s), ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)),
          "profile_read")
    return s.value, a.value, b.value, n.value


def dense_plan(n: int, d: int, dtype: int, nq: int, k: int, path: int = 0) -> dict:
    """The schedule ``ragarc_dense_topk`` would use for this shape (needs the current CUDA device)."""
    out = (ctypes.c_int * 16)()
    check(lib.ragarc_dense_topk_plan(n, d, dtype, nq, k, path, out), "dense_topk_plan")
    keys = ("path", "rows_per_item", "pairs_per_cluster", "query_blocks", "slices", "resident_items",
            "seed_rows", "keep", "tail_slices", "cluster_tiles", "publishing_lists", "published_rank",
            "epilogue_sets")
    return dict(zip(keys, list(out)))


def launch_count() -> int:
    return int(lib.ragarc_launch_count())
