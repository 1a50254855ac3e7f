#!/usr/bin/env python
"""Print an ncu `--metrics gpu__time_duration.sum --csv` launch list as `kernel  ns` lines plus per-kernel totals."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("worm::<unnamed>::", "")
    ns = float(r[vi].replace(",", ""))
    if "-v" in sys.argv:
        print("%-48s %10.0f ns" % (name[:48], ns))
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += ns
all_ns = sum(t[1] for t in tot.values())
for name, (cnt, ns) in tot.items():
    print("%-48s x%-3d %10.1f us  %5.1f %%" % (name[:48], cnt, ns / 1e3, 100 * ns / all_ns))
