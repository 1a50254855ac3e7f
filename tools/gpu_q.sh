#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python tools/pack_case.py 32 4096 400000 0 0 0 | tail -1
python tools/pack_case.py 256 444 100000 0 0 0 | tail -1
python tools/pack_case.py 256 512 100000 0 0 0 | tail -1
python tools/pack_case.py 32 60 400000 0 32 0 | tail -1
python tools/pack_case.py 64 148 200000 0 32 0 | tail -1
