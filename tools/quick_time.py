"""Ad-hoc kernel timing on a B200 (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine

def t_philox(L, n, steps, algo, lanes, reps=3):
    T = np.tile(1.5 + 2.0 * np.arange(64) / 63, (n + 63) // 64)[:n]
    seeds = np.arange(n, dtype=np.uint64)
    st = engine.PhiloxState(L, T, seeds, algo=algo)
    st.run(steps // 10 + 1, lanes_per_chain=lanes)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.run(steps, therm_steps=0, lanes_per_chain=lanes); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    rate = n * steps / best
    print("L=%d n=%d steps=%d algo=%d lanes=%d: %.4f s  %.3e steps/s  (%.1f cycles/step/chain @1.965GHz)" % (
        L, n, steps, algo, lanes, best, rate, 1.965e9 * n / rate), flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for algo in (0, 1):
        for lanes in (1, 4):
            t_philox(32, 4096, 200000, algo, lanes)
    for lanes in (1, 4):
        t_philox(32, 65536, 20000, 0, lanes)
        t_philox(32, 262144, 5000, 1, lanes)
    t_philox(32, 4096, 100000, 0, -1)
    t_philox(64, 512, 100000, 0, 0)
    t_philox(128, 64, 100000, 0, 0)
    t_philox(256, 64, 100000, 0, 0)
