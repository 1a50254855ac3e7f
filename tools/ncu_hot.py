#!/usr/bin/env python
"""Top sampled instructions of one kernel in an .ncu-rep (source page), with addresses and stall reasons.

    python tools/ncu_hot.py gpurun_out/prof_hist.ncu-rep worm_lat [n]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
si, src, ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]


def iv(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot = sum(iv(r[si]) for r in data) or 1
print("%d samples" % tot)
base = int(data[0][0], 16) if data else 0
for r in sorted(data, key=lambda r: -iv(r[si]))[:n]:
    why = sorted(((iv(r[i]), h[6:]) for i, h in stall_cols), reverse=True)[:2]
    print("%6x %6d samples (%4.1f %%) exec %9s  %-60s %s" % (int(r[0], 16) - base, iv(r[si]), 100.0 * iv(r[si]) / tot, r[ex], r[src][:60],
                                                       " ".join("%s=%d" % (h, v) for v, h in why if v)))
