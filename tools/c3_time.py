"""BASELINE configs[2] timing alone: bench.py's c3 section (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from worm_algorithm_b200 import engine
dev = torch.device("cuda", 0)
for r in bench.bench_c3(torch, engine, dev):
    print(json.dumps({k: r[k] for k in ("L", "chains", "dumps_kept", "sim_ms", "worm_steps_per_s", "pipeline_configs_per_s")}))
