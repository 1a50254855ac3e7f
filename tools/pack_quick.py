"""Quick latency / throughput probes of the packed kernel (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.quick_time import t_philox
for algo in (0, 1):
    t_philox(32, 4096, 200000, algo, -2)
    t_philox(32, 33152, 100000, algo, -2)
    t_philox(32, 125000, 40000, algo, -2)
t_philox(32, 125000, 20000, 0, -2, hist=True)
t_philox(64, 8288, 100000, 0, -2)
t_philox(128, 2072, 100000, 0, -2)
