#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt ["note"]
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled", "sm__pipe_tensor_cycles_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu summary of %s\n# %s\n" % (rep, note))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("\n== kernel: %s  grid %s block %s\n" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
            for h, u, v in zip(hdr, units, r):
                if any(h.startswith(k) for k in KEYS):
                    if "stalled" in h and ("pcsamp" in h):
                        continue
                    try:
                        if float(v.replace(",", "")) == 0.0 and "stalled" in h:
                            continue
                    except ValueError:
                        pass
                    f.write("%-90s %-12s %s\n" % (h, u, v))
    print("wrote", out)


if __name__ == "__main__":
    main()
