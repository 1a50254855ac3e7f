#!/bin/bash
# ncu source-level capture of the packed kernel with ONE warp per SM (the latency profile of a lone warp)
mkdir -p gpurun_out
CASE="${1:-32 4096 50000 0 -2 0}"
python tools/pack_case.py $CASE > gpurun_out/plain_pack_lone.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:worm_pack -s 1 -c 1 -f -o gpurun_out/prof_pack_lone python tools/pack_case.py $CASE > gpurun_out/ncu_pack_lone.log 2>&1
tail -1 gpurun_out/plain_pack_lone.log
