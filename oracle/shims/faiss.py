"""numpy stand-in for the ``faiss`` module (test infrastructure).

Lets the reference's own ``encapsulation/database/vector_db/VectorStore_Faiss.py`` be imported and
executed in a container where FAISS is not installed.  Only the symbols that file touches
(``:2,115-142,153,202,263,381,438,467``) exist; flat indexes only (ivf/hnsw are approximate and
outside the exact-search scope).
"""
import numpy as np

from oracle.dense import IndexFlatIP, IndexFlatL2, normalize_L2  # noqa: F401

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
Index = IndexFlatIP


def write_index(index, path):
    """faiss/impl/index_write.cpp for flat indexes: fourcc, header (d, ntotal, two dummies, is_trained,
    metric_type), then the row matrix as a size-prefixed block of 4-byte units (restated here
    independently of the product's reader, rag_arc_b200/formats.py)."""
    import struct
    x = np.ascontiguousarray(index._x, dtype="<f4")
    ip = type(index).__name__ == "IndexFlatIP"
    with open(path, "wb") as f:
        f.write(b"IxFI" if ip else b"IxF2")
        f.write(struct.pack("<i", index.d))
        f.write(struct.pack("<q", x.shape[0]))
        f.write(struct.pack("<qq", 1 << 20, 1 << 20))
        f.write(struct.pack("<B", 1))
        f.write(struct.pack("<i", 0 if ip else 1))
        f.write(struct.pack("<Q", x.size))
        f.write(x.tobytes())


def read_index(path):
    import struct
    with open(path, "rb") as f:
        blob = f.read()
    kind = {b"IxFI": IndexFlatIP, b"IxF2": IndexFlatL2}[blob[:4]]
    d, = struct.unpack_from("<i", blob, 4)
    n, = struct.unpack_from("<q", blob, 8)
    off = 4 + 4 + 8 + 16 + 1 + 4
    count, = struct.unpack_from("<Q", blob, off)
    assert count == n * d
    idx = kind(d)
    idx.add(np.frombuffer(blob, dtype="<f4", count=count, offset=off + 8).reshape(n, d).copy())
    return idx
