#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_physics.py -x -q -k "philox or nibble or hist or G_r or lat_and_wide or segment" 2>&1 | tail -5
python - <<'PY'
import sys
sys.path.insert(0, '.')
from tools.quick_time import t_philox
for L, n in ((32, 4096), (16, 4096), (64, 4096), (128, 1024)):
    t_philox(L, n, 200000 if L <= 32 else 100000, 0, 0, hist=False)
    t_philox(L, n, 200000 if L <= 32 else 100000, 0, 0, hist=True)
t_philox(32, 4096, 200000, 1, 0, hist=True)
t_philox(32, 16384, 50000, 0, 0, hist=True)
t_philox(32, 16384, 50000, 0, 2, hist=True)
t_philox(32, 16384, 50000, 0, -2, hist=True)
PY
