"""Step rate of a bond_flag = 1 run (dump rule of worm_ising_2d.cpp:113-125 every 50 steps, 64 dumps kept per
chain) next to the same run without dumps (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine


def t(L, n, steps, dump_every, max_dumps, nT=64, lanes=0):
    t_idx = np.arange(n) % nT
    T = (1.5 + 2.0 * np.arange(nT) / max(nT - 1, 1))[t_idx]
    seeds = (np.arange(n) // nT).astype(np.uint64)
    # (the first launch of a kernel instantiation loads its module: warm it up on a throw-away state)
    engine.PhiloxState(L, T, seeds, algo=0, group_index=t_idx, n_groups=nT, max_dumps=max_dumps).run(
        2000, therm_steps=200, dump_every=dump_every, lanes_per_chain=lanes)
    st = engine.PhiloxState(L, T, seeds, algo=0, group_index=t_idx, n_groups=nT, max_dumps=max_dumps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); st.run(steps, therm_steps=steps // 10, dump_every=dump_every, lanes_per_chain=lanes); e1.record(); torch.cuda.synchronize()
    s = e0.elapsed_time(e1) * 1e-3
    print("L=%3d n=%6d lanes=%2d steps=%7d dump_every=%3d max_dumps=%3d: %.4f s  %.3e steps/s  dumps/chain %.1f" % (
        L, n, lanes, steps, dump_every, max_dumps, s, n * steps / s, float(st.n_dumps.float().mean()) if max_dumps else 0.0), flush=True)


if __name__ == "__main__":
    for L, n in ((32, 4096), (64, 4096), (128, 1024)):
        t(L, n, 400000, 0, 0)
        t(L, n, 400000, 50, 64)
        t(L, n, 400000, 50, 1024 if L == 32 else 256)
    # large batches run on the packed kernel, whose step range is cut by uniform loop control only
    for L, n in ((32, 33152), (64, 8288)):
        for lanes in (-2,):
            t(L, n, 100000, 0, 0, lanes=lanes)
            t(L, n, 100000, 50, 64, lanes=lanes)
