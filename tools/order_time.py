"""Does the order of the temperatures along the chain range matter? 125 000 chains of L = 32, the bench's 64 temperatures:
temperature-major against one temperature per lane (replica-major, WormSimulation's numbering), with and without G(r)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine
import bench
T64 = bench.sweep_temperatures()
n, steps = 125_000, 40_000
for hist in (False, True):
    for algo in (0, 1):
        for name, t_idx in (("temperature-major", np.repeat(np.arange(64), -(-n // 64))[:n]), ("replica-major", np.arange(n) % 64)):
            st = engine.PhiloxState(32, T64[t_idx], (np.arange(n) // 64).astype(np.uint64), algo=algo, group_index=t_idx.astype(np.int32), n_groups=64, hist=hist)
            st.run(2000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); st.run(steps); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("hist=%d algo=%d %-18s %.2f ms  %.3e steps/s" % (hist, algo, name, ms, n * steps / ms * 1e3), flush=True)
