#!/bin/bash
# large lattices: chains per block x cycles per step for every latency-mode variant
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/bigL_time.log
import sys, os; sys.path.insert(0, ".")
from tools.quick_time import t_philox
for lanes in (32, 8, 4, 0):
    t_philox(64, 4096, 100000, 0, lanes, reps=1)
for lay in ("dual", "compact"):
    os.environ["WORM_LAT_LAYOUT"] = lay
    print(lay)
    t_philox(64, 4096, 100000, 0, 32, reps=1)
    t_philox(128, 1024, 100000, 0, 32, reps=1, nT=8)
del os.environ["WORM_LAT_LAYOUT"]
t_philox(128, 1024, 100000, 0, 0, reps=1, nT=8)
t_philox(256, 512, 100000, 0, 0, reps=1, nT=8)
t_philox(64, 512, 400000, 0, 0, nT=8, reps=1)
t_philox(128, 256, 400000, 0, 0, nT=8, reps=1)
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
