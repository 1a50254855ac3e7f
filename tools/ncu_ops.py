#!/usr/bin/env python
"""Per-opcode stall samples and executed counts of one kernel from an .ncu-rep (source page)."""
import csv
import subprocess
import sys
from collections import Counter

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
si, src, ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")


def iv(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot = sum(iv(r[si]) for r in data) or 1
c, e = Counter(), Counter()
for r in data:
    toks = r[src].split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
    c[op] += iv(r[si])
    e[op] += iv(r[ex])
print("kernel %s: %d samples, %d warp-instructions" % (kern, tot, sum(e.values())))
for op, v in c.most_common(12):
    print("  %-8s %5.1f %% of samples   executed %d" % (op, 100 * v / tot, e[op]))
for r in sorted(data, key=lambda r: -iv(r[si]))[:8]:
    print("  hot: %5s samples  %s" % (r[si], r[src][:80]))
