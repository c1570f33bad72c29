"""Stand-in for the ``rank_bm25`` package (test infrastructure): exposes the oracle restatement
under the name the reference imports at ``core/retrieval/bm25.py:179,402,475``.  The class is
re-declared here so that objects pickled by the reference (``BM25Retriever.save_to_disk``) carry the
module path ``rank_bm25.BM25Okapi`` exactly as they would with the real package."""
from oracle.bm25 import BM25Okapi as _OracleBM25Okapi


class BM25Okapi(_OracleBM25Okapi):
    pass
