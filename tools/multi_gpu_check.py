"""N-rank vs 1-rank equality of the reduced sums (run under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py

Every rank runs its shard of the chains and the sums are all-reduced over NCCL; rank 0 then runs the
whole batch alone and compares the per-temperature accumulators and G(r) histograms bit for bit
(chain ids are global and the random stream is counter-based, so they must be identical)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
from worm_algorithm_b200.worm_simulation import WormSimulation

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
kw = dict(L=32, num_steps=200000, verbose=False, T_arr=1.5 + 2.0 * np.arange(16) / 15.0, bond_flag=0, n_replicas=48,
          rng="philox", hist=True, device=torch.device("cuda", local))
multi = WormSimulation(rank=rank, world_size=world, **kw)
if rank == 0:
    single = WormSimulation(rank=0, world_size=1, **kw)
    ok = bool((multi.results["per_T"] == single.results["per_T"]).all() and (multi.results["hist"] == single.results["hist"]).all())
    same_E = all(multi._E[k] == single._E[k] for k in single._E)
    print("world=%d: per-T sums and G(r) identical to the single-rank run: %s; result dicts identical: %s; Z[T=1.5]=%d" % (
        world, ok, same_E, int(single.results["per_T"][0, 0])), flush=True)
    assert ok and same_E
if world > 1:
    td.barrier()
    td.destroy_process_group()
