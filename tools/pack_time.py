"""Step rates of the packed throughput kernel (worm_pack.cu) next to the other variants (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.quick_time import t_philox

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "l32"):
        for n, steps in ((4096, 200000), (33152, 100000), (125000, 40000), (1000000, 8000)):
            for algo in (0, 1):
                t_philox(32, n, steps, algo, -2)
        t_philox(32, 125000, 40000, 0, 0)
        t_philox(32, 125000, 20000, 0, -2, hist=True)
        t_philox(32, 125000, 20000, 1, -2, hist=True)
        t_philox(32, 125000, 20000, 0, 0, hist=True)
    if which in ("all", "big"):
        t_philox(16, 125000, 40000, 0, -2)
        for L, n in ((64, 4096), (64, 8288), (64, 32768), (128, 1024), (128, 2072), (128, 8192), (256, 512), (256, 444)):
            t_philox(L, n, 100000 if n <= 8288 else 30000, 0, -2)
            t_philox(L, n, 100000 if n <= 8288 else 30000, 1, -2)
        for L, n in ((64, 4096), (128, 1024), (256, 512)):
            t_philox(L, n, 100000, 0, 0)
