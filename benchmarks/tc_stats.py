"""Experiment: counters of the tcgen05 scoring kernel (RAGARC_TC_STATS=1) for one C3-shaped search.
usage: RAGARC_TC_STATS=1 [RAGARC_TC_PUB=0 ...] python benchmarks/tc_stats.py [rows] [nq] [k] [dim]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rag_arc_b200 import _native as N, ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
d = int(sys.argv[4]) if len(sys.argv) > 4 else 768
dev = torch.device("cuda:0")
x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev)
q, _ = synth.dense_queries_cuda(x, nq)
fn = N.lib.ragarc_internal_tc_stats
fn.restype = ctypes.c_int; fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
out = (ctypes.c_ulonglong * 2048)()
for it in range(3):
    ops.dense_topk(x, q, k)
    fn(out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.dense_topk(x, q, k); e1.record(); torch.cuda.synchronize()
fn(out)
names = ["appends", "prunes", "rung_publications", "first_tile_wait_cycles_sum", "first_tile_timeouts", "first_tile_warps"]
print({"env": {k_: v for k_, v in os.environ.items() if k_.startswith("RAGARC_")}, "ms": round(e0.elapsed_time(e1), 4),
       **{nm: int(out[i]) for i, nm in enumerate(names)},
       "plan": N.dense_plan(n, d, N.BF16, nq, k)})

# per-CTA timeline (ns relative to the earliest CTA start): start, first-tile rung published, first-tile wait over, end
import numpy as np
tl = np.array(list(out[16:16 + 320 * 4]), dtype=np.int64).reshape(320, 4)
live = tl[:, 0] > 0
if live.any():
    t0 = tl[live, 0].min()
    rel = np.where(tl > 0, tl - t0, -1)
    for name, rows in (("main", range(0, 160)), ("side", range(160, 320))):
        r = np.array([rel[i] for i in rows if live[i]])
        if len(r) == 0:
            continue
        w = r[:, 2] - r[:, 1]
        print(name, "ctas", len(r), "start us min/med/max", r[:, 0].min() / 1e3, np.median(r[:, 0]) / 1e3, r[:, 0].max() / 1e3,
              "| published us med/max", np.median(r[:, 1]) / 1e3, r[:, 1].max() / 1e3,
              "| wait us med/max", np.median(w) / 1e3, w.max() / 1e3,
              "| end us min/med/max", r[:, 3].min() / 1e3, np.median(r[:, 3]) / 1e3, r[:, 3].max() / 1e3)
        late = [(i, r[i].tolist()) for i in range(len(r)) if r[i, 0] > 5000 or w[i] > 15000]
        print("   late/long-wait CTAs:", late[:24])
