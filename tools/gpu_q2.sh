#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "philox or nibble or segment or dump or events" 2>&1 | tail -3
python tools/dump_time.py 2>&1 | head -9
python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'cyc', d['roofline']['cycles_per_step_per_chain'])
for s in d['lattice_shapes']: print(s)
print(d['c3'])
"
