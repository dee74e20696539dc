#!/bin/bash
# accuracy (drift against the float64 oracle over 300 steps) and throughput as a function of the device solver tolerance
for tol in 3e-6 1e-5 3e-5 1e-4; do
  python scripts/drift_report.py --steps 300 --tol $tol --json gpurun_out/drift_tol_$tol.json > /dev/null 2>&1
  python - $tol <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/drift_tol_{sys.argv[1]}.json'))
for k,v in d.items():
    if isinstance(v,list): print(sys.argv[1], k, 'dq %.2e dv %.2e relF %.2e dr %.2e' % (max(x['max_dqpos'] for x in v), max(x['max_dqvel'] for x in v), max(x['max_rel_dFz'] for x in v), max(x['max_dreward'] for x in v)))
PY
  python bench.py --steps 100 --warmup 20 --no-cpu --tol $tol 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tol value', int(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'iters', round(d['config']['mean_solver_iters'],2))"
done
