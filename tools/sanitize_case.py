"""Smallest run that exercises every production kernel variant (a compute-sanitizer target where that tool is available;
it is closed on the round-2 GPU pool -- there this is just a quick all-variants smoke run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from worm_algorithm_b200 import engine
L, n = 8, 37
T = np.linspace(1.2, 3.4, n); seeds = np.arange(n, dtype=np.uint64) + 11
for lanes in (2, 4, 8, 32, 1, -1, -2):
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
# round 2: packed kernel (both chains, histogram, dumps), nibble layout of the latency kernel, dump intervals of every shape
for algo in (0, 1):
    for L2, n2 in ((16, 70), (64, 40)):
        st = engine.PhiloxState(L2, np.linspace(1.2, 3.4, n2), np.arange(n2, dtype=np.uint64) + 5, algo=algo,
                                group_index=np.arange(n2) % 3, n_groups=3, hist=True, max_dumps=3)
        st.run(260, therm_steps=40, dump_every=50, lanes_per_chain=-2)
        torch.cuda.synchronize()
        print("pack", algo, L2, int(st.acc[:, 0].sum()), int(st.hist.sum()))
for L2, lanes, hist in ((64, 2, False), (64, 2, True), (128, 8, True), (256, 32, False)):
    n2 = 36
    st = engine.PhiloxState(L2, np.linspace(1.2, 3.4, n2), np.arange(n2, dtype=np.uint64) + 5, algo=0,
                            group_index=np.arange(n2) % 3, n_groups=3, hist=hist, max_dumps=2)
    st.run(300, therm_steps=40, dump_every=50, lanes_per_chain=lanes)
    torch.cuda.synchronize()
    print("nib", L2, lanes, int(st.acc[:, 0].sum()))
for de in (2, 10, 16, 48, 50, 100, 7):
    for hist in (False, True):
        st = engine.PhiloxState(8, T, seeds, algo=0, group_index=np.arange(n) % 3, n_groups=3, hist=hist, max_dumps=40)
        st.run(333, therm_steps=37, dump_every=de, lanes_per_chain=2).run(250, therm_steps=37, dump_every=de, lanes_per_chain=2)
        torch.cuda.synchronize()
    print("dump interval", de, int(st.n_dumps.sum()))
dumps = (torch.rand(64, 2 * 32 * 32, device="cuda") < 0.3).to(torch.uint8)
counts = engine.pipeline(dumps, 32)
torch.cuda.synchronize(); print("pipeline", int(counts.sum()))
r = engine.run_mt(8, [2.269, 1.7], [42.0, 7.0], 5000, bond_flag=1, max_events=64, max_trace=16)
img = engine.image(r.dumps[:, :4], 8); blk = engine.block(img, 2, variant=1); cnt = engine.count(img)
torch.cuda.synchronize(); print("mt", r.out[:, :3].tolist(), int(cnt.sum()))
