#!/bin/bash
WORM_B200_SO=$PWD/worm_algorithm_b200/libworm_b200_prof.so timeout 600 python tools/dump_prof.py > gpurun_out/dump_prof.log 2>&1
grep -B1 "max_dumps= 64\|every=  0" gpurun_out/dump_prof.log | grep "lat prof" | head -8
