#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -k "config3 or config2 or substeps or determinism or sharding" 2>&1 | tail -4
bash scripts/variants.sh noWS
for k in 1 2; do python bench.py --steps 150 --warmup 30 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WS value', int(d['value']), 'e2e', int(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'iters', round(d['config']['mean_solver_iters'],2))"; done
for tol in 3e-5 1e-4; do python scripts/parity_report.py --only config3_soft --tol $tol --out gpurun_out/parity_tol_$tol.json 2>&1 | tail -1 | cut -c1-900; done
