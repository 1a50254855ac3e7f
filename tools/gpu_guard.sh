#!/bin/bash
for n in 2 4 8 16; do echo "guard every $n"; WORM_B200_SO=$PWD/worm_algorithm_b200/libworm_b200_g$n.so timeout 300 python tools/dump_prof2.py 2>&1 | grep "lat prof" | tail -2 | awk '{print $4, $6}'; done
