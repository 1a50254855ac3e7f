#!/usr/bin/env python
"""profiles/traffic.json: DRAM bytes (read + write), executed instructions and issue utilisation per launch of the
kernels in an .ncu-rep."""
import csv, json, os, subprocess, sys
rep = sys.argv[1]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
d = {}
try:
    d = json.load(open(out))
except (OSError, ValueError):
    pass
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[2:]:
    rec = dict(zip(h, r)); u = dict(zip(h, units))
    name = rec["Kernel Name"]
    key = "worm_lat_kernel" if "worm_lat_kernel" in name else "worm_wide_kernel" if "worm_wide" in name else name.split("(")[0]
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(rec[m].replace(",", "")) * scale.get(u[m], 1)
    d[key] = tot
    d[key + "_source"] = os.path.basename(rep)
    # executed instructions of the same launch (warp- and thread-level) and the SM's issue utilisation
    for m, k in (("inst_executed", "_warp_inst"), ("sass__thread_inst_executed_true_per_opcode", "_thread_inst"),
                 ("sm__inst_executed.avg.pct_of_peak_sustained_elapsed", "_issue_pct")):
        if m in rec and rec[m] not in ("", "no data"):
            d[key + k] = float(rec[m].replace(",", ""))
json.dump(d, open(out, "w"), indent=1)
print(d)
