"""Ad-hoc kernel timing on a B200 (development aid; bench.py is the contract)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine


def t_philox(L, n, steps, algo=0, lanes=0, hist=False, reps=2, nT=64):
    t_idx = np.arange(n) % nT
    T = (1.5 + 2.0 * np.arange(nT) / max(nT - 1, 1))[t_idx]
    seeds = (np.arange(n) // nT).astype(np.uint64)
    st = engine.PhiloxState(L, T, seeds, algo=algo, group_index=t_idx, n_groups=nT, hist=hist)
    st.run(max(steps // 10, 1), lanes_per_chain=lanes)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.run(steps, therm_steps=0, lanes_per_chain=lanes); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    rate = n * steps / best
    print("L=%3d n=%7d steps=%8d algo=%d lanes=%2d hist=%d: %.4f s  %.3e steps/s  (%.1f cycles/step/chain)" % (
        L, n, steps, algo, lanes, hist, best, rate, 1.965e9 * n / rate), flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "lat"):
        for lanes in (2, 4, 8, 32, 0):
            t_philox(32, 4096, 400000, 0, lanes)
        t_philox(32, 4096, 400000, 1, 0)
        t_philox(32, 4096, 400000, 0, 0, hist=True)
        t_philox(16, 4096, 400000, 0, 0)
        t_philox(64, 512, 400000, 0, 0, nT=8)
        t_philox(128, 256, 400000, 0, 0, nT=8)
        t_philox(256, 128, 400000, 0, 0, nT=8)
        t_philox(32, 8192, 200000, 0, 0)
        t_philox(32, 16384, 100000, 0, 0)
    if which in ("all", "wide"):
        for n in (32768, 131072, 1048576):
            for algo in (0, 1):
                t_philox(32, n, max(4000000000 // n // 8, 1000), algo, 1)
        t_philox(32, 131072, 4000, 0, 0)
        t_philox(32, 131072, 4000, 0, 1, hist=True)
