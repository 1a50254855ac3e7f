#!/bin/bash
# ncu capture (source page) of the solo kernel under the dump rule
mkdir -p gpurun_out
python tools/dump_case.py > gpurun_out/plain_dump.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:worm_lat -s 1 -c 1 -f -o gpurun_out/prof_dump python tools/dump_case.py > gpurun_out/ncu_dump.log 2>&1
tail -2 gpurun_out/ncu_dump.log
