#!/bin/bash
python tools/c5_slice_time.py
python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c5', d['c5']['worm_steps_per_s'], d['c5']['hist_checks'])
for s in d['lattice_shapes'][2:6]: print(s['L'], s['chains'], s['algo'], s['G_r_histograms'], '%.3e'%s['worm_steps_per_s'])
"
