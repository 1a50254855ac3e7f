#!/bin/bash
# solo kernel: parity of the Philox variants, then the step rate with and without G(r)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "philox" 2>&1 | tail -3
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/lat_quick.log
import sys; sys.path.insert(0, ".")
from tools.quick_time import t_philox
t_philox(32, 4096, 400000, 0, 0, hist=False)
t_philox(32, 4096, 400000, 0, 0, hist=True)
t_philox(32, 4096, 400000, 1, 0, hist=False)
t_philox(64, 4096, 100000, 0, 0, reps=1)
t_philox(128, 1024, 100000, 0, 0, reps=1, nT=8)
PY
