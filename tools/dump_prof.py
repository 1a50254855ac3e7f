"""Where a bond_flag = 1 run of the latency kernel spends its cycles (needs the WORM_LAT_PROF variant build:
SRC=worm_lat tools/build_variant.sh prof -DWORM_LAT_PROF=1; WORM_B200_SO=worm_algorithm_b200/libworm_b200_prof.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dump_time import t
import torch

if __name__ == "__main__":
    t(32, 4096, 100000, 0, 0)            # warm-up (module load)
    for L, n in ((32, 4096), (64, 4096), (128, 1024)):
        for de, md in ((0, 0), (50, 0), (50, 64), (64, 64), (48, 64)):
            try:
                t(L, n, 400000, de, md)
            except Exception as e:
                print("failed", L, n, de, md, e)
            torch.cuda.synchronize()
