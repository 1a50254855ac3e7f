#!/bin/bash
# parity tests + bench summary + ncu capture of the worm kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_quick.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("value %.4g  e2e %.4g  cycles/step %.1f  blocking %.0f GB/s" % (d["value"], d["e2e"]["value"], d["roofline"]["cycles_per_step_per_chain"], d["roofline_blocking"]["achieved"]))
PY
bash tools/gpu_ncu.sh > /dev/null 2>&1
ls -la gpurun_out/prof_worm.ncu-rep
