"""GPU: kernel variants that are normally chosen by the planner are forced through their
environment switches (read once per process, hence subprocesses) and must give identical results."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, torch, numpy as np
sys.path.insert(0, %r)
from rag_arc_b200 import ops, synth
from rag_arc_b200.core.retrieval.bm25_index import Bm25Index
dev = torch.device("cuda:0")
x = synth.dense_corpus_cuda(300_000, 256, torch.bfloat16, dev, seed=5)
out = []
for nq, k in ((300, 100), (1000, 50), (40, 10), (1, 5)):
    q, _ = synth.dense_queries_cuda(x, nq, seed=6)
    s, i = ops.dense_topk(x, q, k)
    out.append(i.cpu().numpy()); out.append(s.cpu().numpy())
toks, offs = synth.bm25_corpus_tokens(60_000, vocab=5000, seed=9)
idx = Bm25Index.from_token_ids(toks, offs, device=dev)
qt, ql = idx.encode_query_ids(synth.bm25_queries_tokens(toks, offs, 16, 6, seed=10))
bs, bi = ops.bm25_topk(idx, qt, ql, 20)
out.append(bi.cpu().numpy()); out.append(bs.cpu().numpy())
np.savez(sys.argv[1], *out)
"""


def _run(tmp_path, name, env):
    path = str(tmp_path / f"{name}.npz")
    e = dict(os.environ); e.update(env)
    subprocess.run([sys.executable, "-c", SCRIPT % ROOT, path], check=True, env=e, timeout=600)
    import numpy as np
    z = np.load(path)
    return [z[k] for k in z.files]


def test_forced_variants_agree_bitwise(tmp_path):
    import numpy as np
    base = _run(tmp_path, "default", {})
    for name, env in (("cg1", {"RAGARC_TC_CG": "1"}), ("cg2", {"RAGARC_TC_CG": "2"}),
                      ("cl1", {"RAGARC_TC_CL": "1"}), ("cl2", {"RAGARC_TC_CL": "2"}), ("cl4", {"RAGARC_TC_CL": "4"}),
                      ("slices", {"RAGARC_DENSE_S": "5"}), ("bm25dense", {"RAGARC_BM25_DENSE": "1"})):
        got = _run(tmp_path, name, env)
        for a, b in zip(base, got):
            assert np.array_equal(a, b), name
