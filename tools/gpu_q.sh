#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python tools/pack_case.py 64 4096 200000 0 0 0 | tail -1
python tools/pack_case.py 64 4096 200000 0 0 0 | tail -1
python tools/pack_case.py 64 4096 200000 0 0 1 | tail -1
python tools/pack_case.py 128 1024 200000 0 0 0 | tail -1
python tools/pack_case.py 128 8192 100000 0 0 0 | tail -1
python tools/pack_case.py 256 444 100000 0 0 0 | tail -1
python tools/dump_time.py 2>&1 | head -9 | grep "every= 50 max_dumps= 64\|max_dumps=256"
