"""Host-side cost of a worm_philox_run call (parameter table + upload) on large batches: wall time of a 2-step run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine
import bench
T64 = bench.sweep_temperatures()
for n in (4096, 125_000, 1_000_000):
    t_idx = (np.arange(n) % 64).astype(np.int32)
    st = engine.PhiloxState(32, T64[t_idx], (np.arange(n) // 64).astype(np.uint64), algo=0, group_index=t_idx, n_groups=64, hist=True)
    st.run(2000)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); st.run(2); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.run(2); e1.record(); torch.cuda.synchronize()
    print("n=%8d: 2-step run wall %.2f ms (min of 5), device span %.2f ms" % (n, min(ts) * 1e3, e0.elapsed_time(e1)))
