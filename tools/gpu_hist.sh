#!/bin/bash
# histogram-mode check: parity tests of every worm kernel variant, then the solo kernel's step rate with G(r) on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_physics.py -m gpu -x -q -k "philox or hist or agree" 2>&1 | tail -6
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/hist_time.log
import sys; sys.path.insert(0, ".")
from tools.quick_time import t_philox
t_philox(32, 4096, 400000, 0, 0, hist=True)
t_philox(32, 4096, 400000, 0, 0, hist=False)
t_philox(32, 4736, 400000, 0, 0, hist=True)
t_philox(16, 4736, 400000, 0, 0, hist=True)
t_philox(32, 125000, 20000, 0, 0, hist=True)
PY
