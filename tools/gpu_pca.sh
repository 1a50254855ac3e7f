#!/bin/bash
# PCA row on the GPU box: parity tests, stage timings, ncu launch list (after the plain run succeeded).
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_pca.py -x -q 2>&1 | tail -5
timeout 300 python tools/pca_time.py ${PCA_TIME_ARGS:---no-cpu} > gpurun_out/pca_time.log 2>&1; cat gpurun_out/pca_time.log
if [ "${PCA_NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/pca_launches.csv \
      python tools/pca_time.py --cases 32:20480:10 --reps 1 --no-cpu > gpurun_out/pca_ncu.log 2>&1
  python tools/launch_list.py gpurun_out/pca_launches.csv | head -30
fi
