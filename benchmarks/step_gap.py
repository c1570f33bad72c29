"""Where the C3 step goes beyond its kernels: the per-phase CUDA events inside the library
(ragarc_profile_*) summed over a long back-to-back run against the run's own wall time on the device,
with and without the events, and the scoring phase alone back to back.
usage: python benchmarks/step_gap.py [steps]"""
import sys

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rag_arc_b200 import _native as N
from rag_arc_b200 import ops


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(1)
    x = torch.nn.functional.normalize(torch.randn((1_000_000, 768), generator=g, device=dev), dim=1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn((1024, 768), generator=g, device=dev), dim=1).to(torch.bfloat16)
    step = lambda: ops.dense_topk(x, q, 10)
    for _ in range(6):
        step()
    for rep in range(2):
        plain = timed(step, steps)
        N.profile_enable(True); N.profile_read()
        with_events = timed(step, steps)
        seed, score, merge, nrec = N.profile_read()
        N.profile_enable(False)
        first8 = None
        N.profile_enable(True); N.profile_read()
        torch.cuda.synchronize()
        for _ in range(8):
            step()
        torch.cuda.synchronize()
        s8, sc8, m8, n8 = N.profile_read()
        N.profile_enable(False)
        print(f"rep {rep}: plain {plain:.4f} ms/step | with events {with_events:.4f} ms/step, phases over {nrec} steps: "
              f"setup {seed / nrec:.4f} score {score / nrec:.4f} merge {merge / nrec:.4f} sum {(seed + score + merge) / nrec:.4f} | "
              f"8-step pass: score {sc8 / n8:.4f} merge {m8 / n8:.4f}", flush=True)
    # the dominant kernel alone, back to back: a GEMM of the same shape by cuBLAS for the clock state
    mm = lambda: torch.matmul(q, x.T)
    for _ in range(3):
        mm()
    print(f"cuBLAS bf16 GEMM 1024 x 1M x 768 back to back: {timed(mm, steps):.4f} ms; 8 at a time after idle: {timed(mm, 8):.4f} ms")


if __name__ == "__main__":
    main()
