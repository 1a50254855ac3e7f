#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import sys
sys.path.insert(0, '.')
from tools.quick_time import t_philox
from worm_algorithm_b200 import _lib
for L, n in ((32, 4096), (32, 16384), (64, 4096), (64, 32768), (128, 1024), (128, 2072), (128, 8192), (256, 512), (256, 444), (256, 2048)):
    print("choice", L, n, _lib.philox_choice(L, 0, n, False))
    t_philox(L, n, 100000 if n <= 8288 else 30000, 0, 0)
t_philox(32, 4096, 100000, 0, 0, hist=True)
t_philox(64, 4096, 100000, 0, 0, hist=True)
PY
