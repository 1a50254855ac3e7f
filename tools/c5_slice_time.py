"""A 125 000-chain slice of C5 (what one of eight ranks runs) with G(r): temperature-major (8 temperatures on the rank)
against seed-major numbering (all 64 temperatures on every rank)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine
T64 = np.linspace(1.5, 3.5, 64)
n, steps = 125_000, 20_000
for name, t_idx in (("temperature-major, 8 temperatures", (np.arange(n) // 15_625).astype(np.int32)),
                    ("seed-major, 64 temperatures", (np.arange(n) % 64).astype(np.int32)),
                    ("temperature-major, 64 temperatures", (np.arange(n) * 64 // n).astype(np.int32)),
                    ("super-blocks of 64 temperatures x 125 seeds", ((np.arange(n) % 8000) // 125).astype(np.int32))):
    st = engine.PhiloxState(32, T64[t_idx], (np.arange(n) // 64).astype(np.uint64), algo=0, group_index=t_idx, n_groups=64, hist=True)
    st.run(2000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); st.run(steps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-40s %.2f ms  %.3e steps/s" % (name, ms, n * steps / ms * 1e3))
