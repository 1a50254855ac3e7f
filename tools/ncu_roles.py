#!/usr/bin/env python
"""Stall-reason totals of the walker warp's hot loop in a worm_lat_kernel capture (ncu source page): the instructions
whose execution count equals that of the walker's bond load (LDS.U8 ... executed once per step per walker warp).

    python tools/ncu_roles.py gpurun_out/prof_worm.ncu-rep
"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:worm_lat"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
si, src, ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
def iv(x):
    try: return int(x)
    except ValueError: return 0
# the walker's unrolled round: a long contiguous run of instructions with one execution count that holds bond loads
# (LDS.U8); staging / accountant loops also load bytes, but their runs are short
best = (0, 0, 0)
k = 0
while k < len(data):
    e = iv(data[k][ex]); j = k
    while j + 1 < len(data) and iv(data[j + 1][ex]) == e: j += 1
    if j - k + 1 >= 200 and e > best[0] and any("LDS.U8" in data[m][src] for m in range(k, j + 1)):
        best = (e, k, j)                 # (among the long runs the steady-state loop is the one executed most often)
    k = j + 1
_, lo, hi2 = best
per_slot = iv(data[lo][ex])
sel = data[lo:hi2 + 1]
tot = sum(iv(r[si]) for r in sel)
print("walker hot loop: %d instructions (%.1f KB), executed %d times each, %d samples" % (len(sel), len(sel) * 16 / 1024, per_slot, tot))
agg = collections.Counter()
for r in sel:
    for i, h in stall_cols:
        agg[h[6:]] += iv(r[i])
for h, v in agg.most_common(12):
    print("  %-22s %7d  %5.1f %%" % (h, v, 100.0 * v / max(tot, 1)))
