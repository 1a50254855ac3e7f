#!/bin/bash
for v in prof fake; do echo "variant $v"; WORM_B200_SO=$PWD/worm_algorithm_b200/libworm_b200_$v.so timeout 300 python tools/dump_prof2.py 2>&1 | grep -A1 "lat prof" | tail -6; done
