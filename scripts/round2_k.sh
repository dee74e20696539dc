#!/bin/bash
run() { python bench.py --steps 100 --warmup 10 --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']; print('$*', 'value', int(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'iters', round(c['mean_solver_iters'],2), 'rebuilds', round(c['mean_precond_rebuilds'],2), 'ls', round(c['mean_line_search_evals'],2))"; }
run
run --rebuilds 2
run --rebuilds 4
run --rebuilds 16
run --tol 1e-4
run --envs 8192
run --envs 32768
