#!/bin/bash
# packed kernel: parity first, then step rates
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/pack_time.py ${1:-all} 2>&1 | tee gpurun_out/pack_time.log
