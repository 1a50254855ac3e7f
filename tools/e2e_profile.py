"""Where the end-to-end pass of bench.py spends its time outside the kernel (cProfile of WormSimulation)."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from worm_algorithm_b200.worm_simulation import WormSimulation

temps = 1.5 + 2.0 * np.arange(64) / 63.0
steps = float(sys.argv[1]) if len(sys.argv) > 1 else 1e6


def one():
    return WormSimulation(L=32, run=True, num_steps=steps, verbose=False, T_arr=temps, bond_flag=0, n_replicas=64,
                          rng="philox", algo="taylor", therm_frac=0.1, device="cuda:0", lanes_per_chain=0)


one()
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter()
    one()
    torch.cuda.synchronize()
    print("pass %.1f ms (kernel alone would be %.1f ms)" % ((time.perf_counter() - t0) * 1e3, 4096 * steps / 1.305e11 * 1e3))
pr = cProfile.Profile()
pr.enable()
one()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
