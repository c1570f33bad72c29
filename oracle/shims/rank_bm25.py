"""Stand-in for the ``rank_bm25`` package (test infrastructure): exposes the oracle restatement
under the name the reference imports at ``core/retrieval/bm25.py:179,402,475``."""
from oracle.bm25 import BM25Okapi  # noqa: F401
