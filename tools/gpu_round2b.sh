#!/bin/bash
# Round-2 GPU-box visit, latency kernel only (after a change to worm_lat.cu): parity tests, smoke, bench (+ reference arm),
# launch list, ncu --set full of the two latency-kernel captures.  The other kernels' captures come from tools/gpu_round2.sh.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
NCU_CMD="python bench.py --steps 1 --warmup 3 --num-steps 200000 --no-cpu --e2e-steps 1"
$NCU_CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launches.log 2>&1
$NCU_CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:worm_lat -s 3 -c 1 -f -o gpurun_out/prof_worm $NCU_CMD > gpurun_out/ncu_full.log 2>&1
C3="python tools/pack_case.py 64 4096 50000 0 0 0"
$C3 > gpurun_out/plain_nib.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:worm_lat -s 1 -c 1 -f -o gpurun_out/prof_nib $C3 > gpurun_out/ncu_nib.log 2>&1
# latency kernel with G(r) in shared-memory bins (L = 32) -- summary only (the report would not fit the 64 MiB return limit)
C6="python tools/pack_case.py 32 4096 50000 0 0 1"
$C6 > gpurun_out/plain_sh.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:worm_lat -s 1 -c 1 -f -o /tmp/prof_sh $C6 > gpurun_out/ncu_sh.log 2>&1
python tools/ncu_summary.py /tmp/prof_sh.ncu-rep gpurun_out/r02_sh_ncu_full.txt "round 2, tools/gpu_round2b.sh" > /dev/null 2>&1
python tools/ncu_roles.py /tmp/prof_sh.ncu-rep > gpurun_out/r02_sh_roles.txt 2>&1
python tools/dump_time.py > gpurun_out/dump_time.log 2>&1
ls -la gpurun_out | head -40
du -sh gpurun_out
