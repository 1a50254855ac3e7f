"""Where a 1e6-chain call spends its time: host side of worm_philox_run vs the kernel (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
t_idx = (np.arange(n) // max(n // 64, 1)).clip(max=63).astype(np.int32)
T = (1.5 + 2.0 * np.arange(64) / 63.0)[t_idx]
seeds = (np.arange(n) % 15625).astype(np.uint64)
for hist in (False, True):
    st = engine.PhiloxState(32, T, seeds, algo=0, group_index=t_idx, n_groups=64, hist=hist)
    st.run(2000); torch.cuda.synchronize()
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); st.run(steps); e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print("n=%d hist=%d: host call %.1f ms, events %.1f ms, wall to sync %.1f ms -> %.3e steps/s by events" % (
            n, hist, (t1 - t0) * 1e3, e0.elapsed_time(e1), (t2 - t0) * 1e3, n * steps / e0.elapsed_time(e1) * 1e3), flush=True)
    del st
