"""On-disk formats of an existing RAG-ARC deployment, read without the packages that wrote them.

``FaissVectorStore.save_local`` (/root/reference encapsulation/database/vector_db/
VectorStore_Faiss.py:432-450) leaves two files per store:

* ``<name>.faiss`` - ``faiss.write_index`` of the index.  For the flat indexes on the hot path
  (``IndexFlatIP`` / ``IndexFlatL2``, ``_create_index`` :114-115,125-126,135-136) FAISS's
  serialisation (faiss/impl/index_write.cpp, ``write_index_header`` + ``WRITEXBVECTOR``; unchanged
  since v1.6 for flat indexes apart from the name of the vector) is, little endian::

      char[4]  fourcc        "IxFI" (inner product), "IxF2" (L2), "IxFl" (other metric)
      int32    d
      int64    ntotal
      int64    dummy, dummy  (1 << 20, ignored on read)
      uint8    is_trained
      int32    metric_type   0 = inner product, 1 = L2; > 1: followed by float32 metric_arg
      uint64   count         number of 4-byte units that follow (= ntotal * d)
      float32  xb[count]     the row matrix, row-major

  FAISS is not installable in the build environment, so this layout is restated from the published
  source and pinned by hand-assembled byte strings only (tests/test_formats.py) - "unpinned" against
  a file written by FAISS itself.
* ``<name>.pkl`` - a pickled dict {docstore, index_to_docstore_id, index_type, metric, normalize_L2}
  whose Documents are instances of the REFERENCE's ``core.utils.data_model.Document``; they are
  re-created here as this package's ``Document`` without importing the reference.

``BM25Retriever.save_to_disk`` (/root/reference core/retrieval/bm25.py:550-576) dill-dumps
{vectorizer, docs, k, preprocess_func, bm25_params}; ``load_reference_bm25_state`` reads that file
where neither the reference nor ``rank_bm25`` is importable: the Documents become ours, the
``rank_bm25.BM25Okapi`` object becomes a plain attribute bag (its parameters and idf table are kept
for checking; the postings are rebuilt from the documents), the reference's default tokeniser maps
to ours.  Like the reference's own ``dill.load`` (:596) this trusts the file - a tokeniser pickled by
value is code.
"""
from __future__ import annotations

import io
import pickle
import struct
from typing import Any, Dict, Tuple

import numpy as np

from .core.utils.data_model import Document

_FOURCC_METRIC = {b"IxFI": "ip", b"IxF2": "l2", b"IxFl": None}
_HEADER = struct.Struct("<iqqqBi")           # d, ntotal, dummy, dummy, is_trained, metric_type


def read_faiss_flat(path: str) -> Tuple[np.ndarray, str]:
    """-> (rows float32 [ntotal, d], metric "ip" | "l2").  Raises ValueError for anything that is
    not a flat index (IVF / HNSW files start with other fourccs; approximate indexes are out of scope)."""
    with open(path, "rb") as f:
        fourcc = f.read(4)
        if fourcc not in _FOURCC_METRIC:
            raise ValueError(f"{path}: fourcc {fourcc!r} is not a FAISS flat index (IxFI / IxF2 / IxFl)")
        raw = f.read(_HEADER.size)
        if len(raw) != _HEADER.size:
            raise ValueError(f"{path}: truncated index header")
        d, ntotal, _, _, _trained, metric_type = _HEADER.unpack(raw)
        if metric_type > 1:
            f.read(4)                                            # metric_arg
        if d <= 0 or ntotal < 0:
            raise ValueError(f"{path}: bad header d={d} ntotal={ntotal}")
        (count,) = struct.unpack("<Q", f.read(8))
        if count != ntotal * d:
            raise ValueError(f"{path}: vector block holds {count} floats, header says {ntotal} x {d}")
        rows = np.fromfile(f, dtype="<f4", count=count)
        if rows.size != count:
            raise ValueError(f"{path}: truncated vector block")
    metric = _FOURCC_METRIC[fourcc] or {0: "ip", 1: "l2"}.get(metric_type)
    if metric is None:
        raise ValueError(f"{path}: metric type {metric_type} is neither inner product nor L2")
    return rows.reshape(ntotal, d).astype(np.float32, copy=False), metric


def write_faiss_flat(path: str, rows: np.ndarray, metric: str = "ip") -> None:
    """The inverse of ``read_faiss_flat`` (lets a B200 store hand its fp32 rows back to a FAISS user)."""
    rows = np.ascontiguousarray(rows, dtype="<f4")
    if rows.ndim != 2:
        raise ValueError("rows must be [n, d]")
    if metric not in ("ip", "l2"):
        raise ValueError("metric must be 'ip' or 'l2'")
    n, d = rows.shape
    with open(path, "wb") as f:
        f.write(b"IxFI" if metric == "ip" else b"IxF2")
        f.write(_HEADER.pack(d, n, 1 << 20, 1 << 20, 1, 0 if metric == "ip" else 1))
        f.write(struct.pack("<Q", n * d))
        rows.tofile(f)


class _ReferenceUnpickler(pickle.Unpickler):
    """Maps the reference's value types onto this package's; refuses every other global, so a sidecar
    file cannot run code on load (stock ``pickle.load``, which the reference uses at :462, would)."""

    _ALLOWED = {("core.utils.data_model", "Document"): Document,
                ("utils.data_model", "Document"): Document,
                ("rag_arc_b200.core.utils.data_model", "Document"): Document}

    _SAFE_BUILTINS = {"set": set, "frozenset": frozenset, "bytearray": bytearray, "complex": complex,
                      "range": range, "slice": slice}

    # inert value types that commonly sit in Document.metadata (timestamps, ids, numpy numbers and
    # small arrays, paths): constructing them runs no user code
    _SAFE_VALUES = {
        "datetime": ("datetime", "date", "time", "timedelta", "timezone"),
        "decimal": ("Decimal",), "uuid": ("UUID",), "fractions": ("Fraction",),
        "collections": ("OrderedDict", "defaultdict", "deque", "Counter"),
        "pathlib": ("PurePosixPath", "PosixPath", "PureWindowsPath", "WindowsPath"),
        "numpy": ("dtype", "ndarray", "float32", "float64", "int32", "int64", "bool_", "str_"),
        "numpy.core.multiarray": ("scalar", "_reconstruct"), "numpy._core.multiarray": ("scalar", "_reconstruct"),
        "numpy.core.numeric": ("_frombuffer",), "numpy._core.numeric": ("_frombuffer",),
    }

    def find_class(self, module: str, name: str):
        if module == "builtins" and name in self._SAFE_BUILTINS:      # plain containers inside metadata
            return self._SAFE_BUILTINS[name]
        if name in self._SAFE_VALUES.get(module, ()):
            import importlib
            return getattr(importlib.import_module(module), name)
        try:
            return self._ALLOWED[(module, name)]
        except KeyError:
            raise pickle.UnpicklingError(
                f"sidecar refers to {module}.{name}; only Document objects and plain value types "
                "(datetime, Decimal, UUID, numpy scalars/arrays, paths, collections) are admitted") from None


def load_reference_sidecar(path: str) -> Dict[str, Any]:
    """The ``<name>.pkl`` written by ``FaissVectorStore.save_local`` (or by ``B200VectorStore``):
    dict with ``docstore`` (id -> Document), ``index_to_docstore_id`` (row -> id), ``index_type``,
    ``metric``, ``normalize_L2`` (and ``dtype`` when a B200 store wrote it)."""
    with open(path, "rb") as f:
        data = _ReferenceUnpickler(io.BytesIO(f.read())).load()
    if not isinstance(data, dict) or "docstore" not in data or "index_to_docstore_id" not in data:
        raise ValueError(f"{path}: not a vector-store sidecar")
    for key, doc in data["docstore"].items():
        if not isinstance(doc, Document):
            raise ValueError(f"{path}: docstore entry {key!r} is not a Document")
    return data


class ForeignBM25:
    """What a pickled ``rank_bm25.BM25Okapi`` turns into here: its attributes (``k1``, ``b``,
    ``epsilon``, ``corpus_size``, ``avgdl``, ``idf``, ``doc_len``, ``doc_freqs`` ...) without its code."""


def load_reference_bm25_state(path: str) -> Dict[str, Any]:
    """-> {"vectorizer": ForeignBM25 | object, "docs": [Document], "k", "preprocess_func", "bm25_params"}."""
    import dill
    from .core.retrieval import bm25 as our_bm25

    mapped = {("rank_bm25", "BM25Okapi"): ForeignBM25,
              ("core.utils.data_model", "Document"): Document,
              ("utils.data_model", "Document"): Document,
              ("core.retrieval.bm25", "default_preprocessing_func"): our_bm25.default_preprocessing_func,
              ("retrieval.bm25", "default_preprocessing_func"): our_bm25.default_preprocessing_func}

    class _Unpickler(dill.Unpickler):
        def find_class(self, module, name):
            hit = mapped.get((module, name))
            return hit if hit is not None else super().find_class(module, name)

    with open(path, "rb") as f:
        state = _Unpickler(f).load()
    if not isinstance(state, dict) or not {"vectorizer", "docs", "k"} <= set(state):
        raise ValueError(f"{path}: not a BM25 retriever state")
    return state
