#!/bin/bash
# packed kernel: parity first, then step rates, then an ncu capture of one wave of L = 32 chains
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "philox" 2>&1 | tail -5
timeout 600 python tools/pack_time.py ${1:-all} 2>&1 | tee gpurun_out/pack_time.log
if [ "$2" != "noncu" ]; then
CASE="python tools/pack_case.py 32 33152 20000 0 -2 0"
$CASE > gpurun_out/plain_pack.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:worm_pack -s 1 -c 1 -f -o gpurun_out/prof_pack $CASE > gpurun_out/ncu_pack.log 2>&1
tail -2 gpurun_out/ncu_pack.log
fi
