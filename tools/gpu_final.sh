#!/bin/bash
# last visit of the round: the whole GPU suite, smoke, default bench line
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python tools/mt_time.py > gpurun_out/mt_time.log 2>&1; cat gpurun_out/mt_time.log
