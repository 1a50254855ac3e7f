#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_physics.py tests/test_gpu_tree.py -x -q -k "philox or nibble or segment or dump or tree or bonds or events" 2>&1 | tail -5
python tools/dump_time.py 2>&1 | tail -14
