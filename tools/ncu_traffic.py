#!/usr/bin/env python
"""profiles/traffic.json: DRAM bytes (read + write), executed instructions, issue / pipe utilisation and
shared-memory conflicts per launch of the kernels in an .ncu-rep (`ncu --set full`).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep [chain_steps_of_the_worm_launch]
"""
import csv, json, os, subprocess, sys
rep = sys.argv[1]
chain_steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
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
    key = None
    for k in ("worm_lat_kernel", "worm_wide_kernel", "worm_pack_kernel", "worm_mt_kernel"):
        if k in name:
            key = k
    if key is None:
        key = name.split("(")[0].split("<")[0].split("::")[-1].strip()
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(rec[m].replace(",", "")) * scale.get(u[m], 1)
    d[key] = tot
    d[key + "_source"] = os.path.basename(rep)
    for m, k in (("inst_executed", "_warp_inst"), ("sass__thread_inst_executed_true_per_opcode", "_thread_inst"),
                 ("smsp__thread_inst_executed.sum", "_thread_inst"),
                 ("sm__inst_executed.avg.pct_of_peak_sustained_elapsed", "_issue_pct"),
                 ("smsp__issue_active.avg.pct_of_peak_sustained_active", "_issue_pct"),
                 ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "_alu_pct"),
                 ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "_fmaheavy_pct"),
                 ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "_bank_conflicts"),
                 ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "_smem_wavefronts"),
                 ("gpu__time_duration.sum", "_duration")):
        if m in rec and rec[m] not in ("", "no data", "n/a"):
            try:
                d[key + k] = float(rec[m].replace(",", ""))
            except ValueError:
                pass
    if chain_steps and key.startswith("worm_"):
        d[key + "_chain_steps"] = chain_steps
json.dump(d, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in d.items() if not k.endswith("_note")}, indent=1))
