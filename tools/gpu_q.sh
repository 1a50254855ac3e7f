#!/bin/bash
for v in "" _r2; do echo "variant [$v]"; if [ -n "$v" ]; then export WORM_B200_SO=$PWD/worm_algorithm_b200/libworm_b200$v.so; fi
python tools/pack_case.py 32 4096 800000 0 0 0 | tail -1
python tools/pack_case.py 32 4096 800000 0 0 0 | tail -1
python tools/pack_case.py 64 4096 400000 0 0 0 | tail -1
python tools/pack_case.py 32 4096 400000 0 0 1 | tail -1
done
