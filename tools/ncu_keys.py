#!/usr/bin/env python
"""Compact per-kernel view of an `ncu --page raw --csv` dump: a fixed list of metrics plus PC-sampling stall counts."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("== %s" % d["Kernel Name"][:70])
    for k in KEYS:
        if k in d:
            print("   %-70s %s %s" % (k, d[k], u[k]))
    st = []
    for k in hdr:
        if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k:
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v > 0:
                st.append((v, k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
    tot = sum(v for v, _ in st) or 1
    print("   stalls: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for v, k in sorted(st, reverse=True)[:7]))
