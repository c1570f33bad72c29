#!/usr/bin/env python
"""Per-CUDA-source-line warp-stall samples and executed instructions of one kernel out of an .ncu-rep
(``ncu --page source --print-source cuda,sass``, read without a GPU).
usage: ncu_source_lines.py REP KERNEL_REGEX [top]"""
import csv, os, re, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fpath, func, hdr, lines = None, None, None, []
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = os.path.basename(r[1]); hdr = None; continue
    if r[0] == "Function Name":
        func = r[1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or not re.search(kre, func or "") or not r[0].isdigit():
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        s, i = int(d["# Samples"]), int(d["Instructions Executed"])
    except (KeyError, ValueError):
        continue
    stalls = sorted(((int(v), k[6:]) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)), reverse=True)[:3]
    lines.append((s, i, f"{fpath}:{r[0]}", r[1].strip()[:100], stalls))
ts, ti = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"# {os.path.basename(rep)}: kernel ~ /{kre}/: {ts} warp-stall samples, {ti} warp instructions")
for s, i, where, src, stalls in sorted(lines, reverse=True)[:top]:
    print(f"{100.0*s/max(ts,1):5.1f}% samples {100.0*i/max(ti,1):5.1f}% inst  {where}: {src}   [{' '.join(f'{n}={v}' for v, n in stalls)}]")
