#!/bin/bash
# developer loop: bench every kernel variant built under build/ (USIM_LIB override)
for v in "$@"; do USIM_LIB=$PWD/build/libusim_$v.so python bench.py --steps 150 --warmup 30 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v value', int(d['value']), 'e2e', int(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'iters', round(d['config']['mean_solver_iters'],2))"; done
