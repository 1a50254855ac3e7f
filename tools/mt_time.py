"""Step rate of the deterministic (MT19937) mode: the reference's default sweep, one chain per temperature."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine
for L, nT, steps in ((32, 26, 2_000_000), (32, 1, 2_000_000), (32, 1024, 200_000), (8, 26, 2_000_000)):
    T = np.linspace(1.0, 3.5, nT); seeds = np.arange(nT) + 0.5
    engine.run_mt(L, T, seeds, 1000)
    torch.cuda.synchronize(); t0 = time.time()
    r = engine.run_mt(L, T, seeds, steps)
    torch.cuda.synchronize(); dt = time.time() - t0
    tot = float(r.out[:, 2].sum())
    print("L=%d chains=%d: %.3f s, %.3e steps/s total, %.3e per chain (%.0f cycles/step)" % (L, nT, dt, tot / dt, tot / dt / nT, 1.965e9 * nT * dt / tot))
