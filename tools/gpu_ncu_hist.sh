#!/bin/bash
# ncu capture (source page) of the solo kernel with G(r) histograms
mkdir -p gpurun_out
python tools/hist_case.py > gpurun_out/plain_hist.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:worm_lat -s 1 -c 1 -f -o gpurun_out/prof_hist python tools/hist_case.py > gpurun_out/ncu_hist.log 2>&1
tail -2 gpurun_out/ncu_hist.log
