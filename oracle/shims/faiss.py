"""numpy stand-in for the ``faiss`` module (test infrastructure).

Lets the reference's own ``encapsulation/database/vector_db/VectorStore_Faiss.py`` be imported and
executed in a container where FAISS is not installed.  Only the symbols that file touches
(``:2,115-142,153,202,263,381,438,467``) exist; flat indexes only (ivf/hnsw are approximate and
outside the exact-search scope).
"""
import pickle

import numpy as np

from oracle.dense import IndexFlatIP, IndexFlatL2, normalize_L2  # noqa: F401

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
Index = IndexFlatIP


def write_index(index, path):
    with open(path, "wb") as f:
        pickle.dump({"kind": type(index).__name__, "d": index.d, "x": index._x}, f)


def read_index(path):
    with open(path, "rb") as f:
        blob = pickle.load(f)
    idx = {"IndexFlatIP": IndexFlatIP, "IndexFlatL2": IndexFlatL2}[blob["kind"]](blob["d"])
    idx.add(blob["x"])
    return idx
