#!/bin/bash
# two B200s: the 2-rank pytest case and the N = 2 bench line (torchrun, NCCL, 127.0.0.1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_physics.py -x -q -m gpu -k "multi_gpu" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -2 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.4e e2e %.4e' % (d['value'], d['e2e']['value']), d['check']['multi_gpu_equal'], 'c5 %.3e' % d['c5']['worm_steps_per_s'], d['clocks'])
PY
