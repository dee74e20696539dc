#!/bin/bash
# developer loop: GPU parity suite + three short bench runs (prints throughput, kernel time, mean CG iterations)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for k in 1 2 3; do python bench.py --steps 150 --warmup 30 --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', int(d['value']), 'e2e', int(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'iters', round(d['config']['mean_solver_iters'],2))"; done
