"""``NativeFlatIndex``: the library-owned flat index (``ragarc_index_*`` in include/ragarc_b200.h)
driven with plain numpy host buffers - no torch tensor on the path.

This is what a host that is not PyTorch sees: create, ``add`` fp32 rows, ``search`` fp32 queries,
``remove`` rows, exactly the calls ``FaissVectorStore`` makes on ``faiss.IndexFlatIP``
(VectorStore_Faiss.py:114-115 create, :178,202 normalise + add, :259-263 normalise + search,
:385-419 remove_ids).  Device memory, workspace and staging live inside the handle.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as N

_DTYPES = {"float32": N.F32, "bfloat16": N.BF16, "float16": N.F16}
_METRICS = {"ip": N.METRIC_IP, "cosine": N.METRIC_COSINE, "l2": N.METRIC_L2}


def pinned_array(shape, dtype) -> np.ndarray:
    """numpy array over page-locked memory from ``ragarc_host_alloc`` (freed with the array): queries
    and results handed to ``NativeFlatIndex.search(..., out=)`` in such arrays travel by direct DMA."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = ctypes.c_void_p()
    N.check(N.lib.ragarc_host_alloc(nbytes, ctypes.byref(p)), "host_alloc")
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[id(buf)] = _Pinned(p, buf)
    arr_base = arr
    import weakref
    weakref.finalize(arr_base, _PINNED.pop, id(buf), None)
    return arr


class _Pinned:
    def __init__(self, p, buf):
        self.p, self.buf = p, buf

    def __del__(self):
        try:
            N.lib.ragarc_host_free(self.p)
        except Exception:
            pass


_PINNED: dict = {}


class NativeFlatIndex:
    def __init__(self, d: int, dtype: str = "float32", metric: str = "cosine"):
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        if metric not in _METRICS:
            raise ValueError(f"metric must be one of {sorted(_METRICS)}")
        self.d, self.dtype, self.metric = int(d), dtype, metric
        h = ctypes.c_void_p()
        N.check(N.lib.ragarc_index_create(self.d, _DTYPES[dtype], _METRICS[metric], ctypes.byref(h)), "index_create")
        self._h = h

    # faiss.IndexFlatIP.ntotal
    @property
    def ntotal(self) -> int:
        return int(N.lib.ragarc_index_ntotal(self._h))

    def reserve(self, capacity: int) -> None:
        N.check(N.lib.ragarc_index_reserve(self._h, int(capacity), None), "index_reserve")

    def add(self, x: np.ndarray) -> None:
        """x: float32 [n, d] host array (normalised inside when the metric is cosine)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"expected [n,{self.d}] float32 rows")
        N.check(N.lib.ragarc_index_add(self._h, x.ctypes.data, x.shape[0], 1, None), "index_add")

    def search(self, q: np.ndarray, k: int, out=None):
        """q: float32 [nq, d] host array -> (D float32 [nq,k] descending, I int64 [nq,k], -1 padded):
        the return contract of ``faiss.IndexFlatIP.search`` (metric "l2": squared distances ascending,
        that of ``faiss.IndexFlatL2.search``)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.d:
            raise ValueError(f"expected [nq,{self.d}] float32 queries")
        if out is not None:
            D, I = out
            if D.shape != (q.shape[0], k) or I.shape != D.shape or D.dtype != np.float32 or I.dtype != np.int64:
                raise ValueError("out must be (float32 [nq,k], int64 [nq,k])")
        else:
            D = np.empty((q.shape[0], k), np.float32)
            I = np.empty((q.shape[0], k), np.int64)
        N.check(N.lib.ragarc_index_search(self._h, q.ctypes.data, q.shape[0], int(k), D.ctypes.data,
                                          I.ctypes.data, 1, None), "index_search")
        return D, I

    def remove(self, rows) -> int:
        """Drops the given row numbers; survivors are renumbered densely in order (remove_ids)."""
        rows = np.ascontiguousarray(rows, dtype=np.int64).ravel()
        before = self.ntotal
        N.check(N.lib.ragarc_index_remove(self._h, rows.ctypes.data, rows.shape[0], None), "index_remove")
        return before - self.ntotal

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            N.lib.ragarc_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class NativeShardedIndex:
    """``ragarc_sharded_*``: the same index row-sharded over several GPUs (or several shards on one)
    and driven by this one process - numpy host buffers in and out, results identical to
    ``NativeFlatIndex`` for any number of shards."""

    def __init__(self, d: int, dtype: str = "float32", metric: str = "cosine", devices=(0,)):
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        if metric not in _METRICS:
            raise ValueError(f"metric must be one of {sorted(_METRICS)}")
        self.d, self.dtype, self.metric, self.devices = int(d), dtype, metric, tuple(int(x) for x in devices)
        devs = (ctypes.c_int * len(self.devices))(*self.devices)
        h = ctypes.c_void_p()
        N.check(N.lib.ragarc_sharded_create(self.d, _DTYPES[dtype], _METRICS[metric], len(self.devices), devs,
                                            ctypes.byref(h)), "sharded_create")
        self._h = h

    @property
    def ntotal(self) -> int:
        return int(N.lib.ragarc_sharded_ntotal(self._h))

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise ValueError(f"expected [n,{self.d}] float32 rows")
        N.check(N.lib.ragarc_sharded_add(self._h, x.ctypes.data, x.shape[0]), "sharded_add")

    def search(self, q: np.ndarray, k: int, out=None):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.d:
            raise ValueError(f"expected [nq,{self.d}] float32 queries")
        D = np.empty((q.shape[0], k), np.float32)
        I = np.empty((q.shape[0], k), np.int64)
        N.check(N.lib.ragarc_sharded_search(self._h, q.ctypes.data, q.shape[0], int(k), D.ctypes.data, I.ctypes.data),
                "sharded_search")
        return D, I

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            N.lib.ragarc_sharded_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass
