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
r = engine.run_mt(8, [2.269, 1.7], [42.0, 7.0], 5000, bond_flag=1, max_events=64, max_trace=16)
img = engine.image(r.dumps[:, :4], 8); blk = engine.block(img, 2, variant=1); cnt = engine.count(img)
torch.cuda.synchronize(); print("mt", r.out[:, :3].tolist(), int(cnt.sum()))
