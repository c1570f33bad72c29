"""Where the time of one synchronous ragarc_index_search call goes (C3 shape): device pointers vs
pinned host buffers vs pageable host buffers, and the kernel phases reported by the library."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rag_arc_b200 import _native as N, synth
from rag_arc_b200.native_index import pinned_array
dev = torch.device("cuda:0")
n, d, nq, k = 1_000_000, 768, 1024, 100
x = synth.dense_corpus_cuda(n, d, torch.bfloat16, dev)
q = torch.nn.functional.normalize(torch.randn((nq, d), device=dev), dim=1)
h = ctypes.c_void_p()
N.check(N.lib.ragarc_index_create(d, N.BF16, N.METRIC_COSINE, ctypes.byref(h)), "create")
N.check(N.lib.ragarc_index_reserve(h, n, None), "reserve")
for a in range(0, n, 131072):
    c = x[a:a + 131072].float()
    N.check(N.lib.ragarc_index_add(h, c.data_ptr(), c.shape[0], 0, None), "add")
torch.cuda.synchronize()
Dd = torch.empty((nq, k), dtype=torch.float32, device=dev); Id = torch.empty((nq, k), dtype=torch.int64, device=dev)
qp = pinned_array((nq, d), np.float32); qp[:] = q.cpu().numpy()
Dp = pinned_array((nq, k), np.float32); Ip = pinned_array((nq, k), np.int64)
qg = np.array(qp, copy=True); Dg = np.empty((nq, k), np.float32); Ig = np.empty((nq, k), np.int64)
def timeit(fn, n_it=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n_it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3 / n_it
def dev_call():
    N.check(N.lib.ragarc_index_search(h, q.data_ptr(), nq, k, Dd.data_ptr(), Id.data_ptr(), 0, None), "s"); torch.cuda.synchronize()
def pin_call():
    N.check(N.lib.ragarc_index_search(h, qp.ctypes.data, nq, k, Dp.ctypes.data, Ip.ctypes.data, 1, None), "s")
def page_call():
    N.check(N.lib.ragarc_index_search(h, qg.ctypes.data, nq, k, Dg.ctypes.data, Ig.ctypes.data, 1, None), "s")
st = torch.cuda.Stream(dev)
def dev_call_stream():
    N.check(N.lib.ragarc_index_search(h, q.data_ptr(), nq, k, Dd.data_ptr(), Id.data_ptr(), 0, st.cuda_stream), "s"); st.synchronize()
def pin_call_stream():
    N.check(N.lib.ragarc_index_search(h, qp.ctypes.data, nq, k, Dp.ctypes.data, Ip.ctypes.data, 1, st.cuda_stream), "s")
res = {"device_ptrs_ms": timeit(dev_call), "pinned_ms": timeit(pin_call), "pageable_ms": timeit(page_call),
       "device_ptrs_own_stream_ms": timeit(dev_call_stream), "pinned_own_stream_ms": timeit(pin_call_stream)}
N.profile_enable(True); N.profile_read()
for _ in range(10): pin_call()
seed, score, merge, cnt = N.profile_read(); N.profile_enable(False)
res.update({"kernel_score_ms": score / cnt, "kernel_merge_ms": merge / cnt, "setup_ms": seed / cnt})
qt = torch.from_numpy(np.array(qp)).pin_memory()
def h2d(): qt.to(dev, non_blocking=True); torch.cuda.synchronize()
def d2h(): Dd.cpu(); Id.cpu()
res.update({"h2d_3MB_ms": timeit(h2d), "d2h_1.2MB_pageable_ms": timeit(d2h)})
print(res)
