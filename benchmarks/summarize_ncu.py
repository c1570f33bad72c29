#!/usr/bin/env python
"""Turn an .ncu-rep (read here, without a GPU) into the small text/JSON summaries kept under
profiles/: key raw metrics per kernel, details-page sections, and the top stall lines."""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        rec = {"kernel": d.get("Kernel Name", "")[:80]}
        for k in KEYS:
            for h in hdr:
                if h == k or h.endswith("." + k):
                    rec[k] = {"value": d[h], "unit": units[hdr.index(h)]}
        res.append(rec)
    return res


def main():
    rep, out = sys.argv[1], sys.argv[2]
    recs = raw(rep)
    with open(out, "w") as f:
        json.dump(recs, f, indent=1)
    for r in recs:
        print(r["kernel"])
        for k, v in r.items():
            if k != "kernel":
                print("   ", k, v["value"], v["unit"])


if __name__ == "__main__":
    main()
