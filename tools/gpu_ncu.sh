#!/bin/bash
# ncu captures of the bench's worm kernel (short chains so the ~40 replays stay cheap)
set -x
mkdir -p gpurun_out
NCU_CMD="python bench.py --steps 1 --warmup 3 --num-steps 200000 --no-cpu --e2e-steps 1"
$NCU_CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:worm_ -s 3 -c 1 -o gpurun_out/prof_worm $NCU_CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
