#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "random_batches" 2>&1 | tail -4
