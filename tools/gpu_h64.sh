#!/bin/bash
C="python tools/pack_case.py 64 4096 50000 0 0 1"
$C > gpurun_out/plain_h64.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:worm_lat -s 1 -c 1 -f -o gpurun_out/prof_h64 $C > gpurun_out/ncu_h64.log 2>&1
tail -1 gpurun_out/plain_h64.log
