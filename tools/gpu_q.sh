#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "philox or nibble or segment or dump or events or hist or random" 2>&1 | tail -3
python tools/pack_case.py 32 4096 400000 0 0 0 | tail -1
python tools/pack_case.py 32 4096 400000 0 0 0 | tail -1
python tools/pack_case.py 32 4096 400000 0 0 1 | tail -1
python tools/pack_case.py 64 4096 200000 0 0 0 | tail -1
python tools/pack_case.py 64 4096 200000 0 0 1 | tail -1
python tools/pack_case.py 128 1024 200000 0 0 0 | tail -1
python tools/dump_time.py 2>&1 | head -9 | grep "every= 50 max_dumps= 64"
