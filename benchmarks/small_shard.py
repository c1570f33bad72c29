"""The per-rank work of a row-sharded C3 search on ONE GPU: 1024 queries against 1M / G rows for
G = 1, 2, 4, 8 - whole step (CUDA-graph replay) and the library's own phase events, against the
tensor roof.  usage: python benchmarks/small_shard.py [steps] [G,G,...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_arc_b200 import _native as N, ops, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
xall = synth.dense_corpus_cuda(1_000_000, 768, torch.bfloat16, dev)
q, _ = synth.dense_queries_cuda(xall, 1024)
for G in ([int(g) for g in sys.argv[2].split(',')] if len(sys.argv) > 2 else (8, 4, 2, 1)):
    n = 1_000_000 // G
    x = xall[:n]
    step = lambda: ops.dense_topk(x, q, 100, n_rows=n)
    for _ in range(4):
        step()
    N.profile_enable(True); N.profile_read()
    torch.cuda.synchronize()
    for _ in range(16):
        step()
    torch.cuda.synchronize()
    seed, score, merge, nrec = N.profile_read(); N.profile_enable(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flop = 2.0 * 1024 * n * 768
    print(json.dumps({"G": G, "rows": n, "step_ms": round(ms, 4), "score_ms": round(score / nrec, 4), "merge_ms": round(merge / nrec, 4),
                      "step_tflops": round(flop / ms / 1e9, 1), "score_tflops": round(flop / (score / nrec) / 1e9, 1),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("RAGARC_TC")},
                      "plan": {k: v for k, v in N.dense_plan(n, 768, N.BF16, 1024, 100).items()
                               if k in ("slices", "resident_items", "tail_slices", "cluster_tiles", "publishing_lists", "published_rank", "keep")}}), flush=True)
