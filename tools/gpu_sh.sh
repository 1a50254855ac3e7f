#!/bin/bash
python tools/pack_case.py 32 4096 400000 0 0 1 | tail -1
python tools/pack_case.py 32 4096 400000 0 0 1 | tail -1
python tools/pack_case.py 16 4096 400000 0 0 1 | tail -1
python tools/pack_case.py 32 4096 400000 0 0 0 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "hist or segment or random" 2>&1 | tail -2
