"""Smallest run that exercises every production kernel variant (for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine
L, n = 8, 37
T = np.linspace(1.2, 3.4, n); seeds = np.arange(n, dtype=np.uint64) + 11
for lanes in (2, 4, 8, 32, 1, -1):
    st = engine.PhiloxState(L, T, seeds, algo=0, group_index=np.arange(n) % 3, n_groups=3, hist=True, max_dumps=8)
    st.run(333, therm_steps=40, dump_every=50, lanes_per_chain=lanes).run(200, therm_steps=40, dump_every=50, lanes_per_chain=lanes)
    torch.cuda.synchronize()
    print(lanes, int(st.acc[:, 0].sum()), int(st.hist.sum()))
# large lattices: the multi-walker variants of the latency kernel (3 walkers x 4 chains, compact 4 x 1)
for L2, lanes in ((64, 8), (64, 4), (128, 32), (32, 2)):
    n2 = 40
    st = engine.PhiloxState(L2, np.linspace(1.2, 3.4, n2), np.arange(n2, dtype=np.uint64) + 5, algo=0,
                            group_index=np.arange(n2) % 3, n_groups=3, hist=True, max_dumps=2)
    st.run(300, therm_steps=40, dump_every=50, lanes_per_chain=lanes)
    torch.cuda.synchronize()
    print(L2, lanes, int(st.acc[:, 0].sum()), int(st.hist.sum()))
r = engine.run_mt(8, [2.269, 1.7], [42.0, 7.0], 5000, bond_flag=1, max_events=64, max_trace=16)
img = engine.image(r.dumps[:, :4], 8); blk = engine.block(img, 2, variant=1); cnt = engine.count(img)
torch.cuda.synchronize(); print("mt", r.out[:, :3].tolist(), int(cnt.sum()))
