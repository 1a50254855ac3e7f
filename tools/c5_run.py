"""BASELINE.json configs[4]: 10^6 independent L = 32 chains (64 temperatures x 15625 seeds) with G(r)
histograms, sharded over the ranks, one NCCL all-reduce of the per-temperature sums and histograms.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_run.py [steps]

Prints worm steps/s of the whole job (wall clock of the public WormSimulation call, max over ranks) and checks
the reduced histogram: bin 0 of every temperature equals its Z, every bin sums to the measured iterations."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
from worm_algorithm_b200.worm_simulation import WormSimulation
from worm_algorithm_b200._lib import ACC_ITERS, ACC_Z

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    td.init_process_group("nccl", device_id=dev)
steps = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000
n_rep = int(sys.argv[2]) if len(sys.argv) > 2 else 15625
temps = 1.5 + 2.0 * np.arange(64) / 63.0
kw = dict(L=32, num_steps=steps, verbose=False, T_arr=temps, bond_flag=0, n_replicas=n_rep, rng="philox", hist=True,
          device=dev, rank=rank, world_size=world)
WormSimulation(**dict(kw, num_steps=2000))          # warm-up: allocations, NCCL communicator
best = None
for _ in range(2):
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim = WormSimulation(**kw)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    best = float(t[0]) if best is None else min(best, float(t[0]))
if rank == 0:
    per_T, hist = sim.results["per_T"].cpu().numpy(), sim.results["hist"].cpu().numpy()
    ok0 = bool((hist[:, 0] == per_T[:, ACC_Z]).all())
    oks = bool((hist.sum(axis=1) == per_T[:, ACC_ITERS]).all())
    n = 64 * n_rep
    print("world=%d: %d chains x %d steps with G(r): %.3f s -> %.3e worm steps/s (%.3e per GPU); hist[:,0] == Z: %s; "
          "sum of bins == measured iterations: %s" % (world, n, steps, best, n * steps / best, n * steps / best / world, ok0, oks), flush=True)
    assert ok0 and oks
if world > 1:
    td.barrier()
    td.destroy_process_group()
