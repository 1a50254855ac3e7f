#!/bin/bash
# ncu --set full captures of the packed kernel: Taylor one wave, binary 125000 chains, Taylor with G(r)
mkdir -p gpurun_out
i=0
for CASE in "32 33152 20000 0 -2 0" "32 125000 10000 1 -2 0" "32 33152 20000 0 -2 1"; do
  i=$((i+1))
  python tools/pack_case.py $CASE > gpurun_out/plain_pack$i.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:worm_pack -s 1 -c 1 -f -o gpurun_out/prof_pack$i python tools/pack_case.py $CASE > gpurun_out/ncu_pack$i.log 2>&1
  tail -1 gpurun_out/plain_pack$i.log
done
