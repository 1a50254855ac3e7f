#!/bin/bash
# quick confirmation after a latency-kernel change: GPU tests, dump-rule timing, walker cycle counters
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/dump_time.py > gpurun_out/dump_time.log 2>&1; cat gpurun_out/dump_time.log
WORM_B200_SO=$PWD/worm_algorithm_b200/libworm_b200_prof.so timeout 600 python tools/dump_prof.py > gpurun_out/dump_prof.log 2>&1
grep -c "lat prof" gpurun_out/dump_prof.log
