#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tree.py -x -q -k "philox or nibble or segment or dump or tree or bonds or events" 2>&1 | tail -5
WORM_B200_SO=$PWD/worm_algorithm_b200/libworm_b200_prof.so timeout 600 python tools/dump_prof.py > gpurun_out/dump_prof.log 2>&1
grep -B1 "max_dumps= 64\|every=  0" gpurun_out/dump_prof.log
