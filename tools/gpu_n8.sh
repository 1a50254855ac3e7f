#!/bin/bash
# eight B200s: the N = 8 bench line (torchrun, NCCL, 127.0.0.1)
mkdir -p gpurun_out
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value %.4e e2e %.4e' % (d['value'], d['e2e']['value']), d['check']['multi_gpu_equal'], 'c5 %.3e' % d['c5']['worm_steps_per_s'], d['clocks'])
PY
