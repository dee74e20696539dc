python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rb in 1 2 3 4 8; do python bench.py --steps 100 --warmup 20 --no-cpu --rebuilds $rb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rebuilds $rb', 'value', int(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'iters', round(d['config']['mean_solver_iters'],2))"; done
