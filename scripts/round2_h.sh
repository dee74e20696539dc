#!/bin/bash
bash scripts/closed_loop_pair.sh r02h profiles/r02_probe_fit2.json 2>&1 | tail -22
python scripts/parity_report.py --out gpurun_out/parity_drift_r02h.json 2>&1 | cut -c1-1000
for tol in 1e-5 1e-4; do python scripts/parity_report.py --only long --tol $tol --out gpurun_out/parity_long_$tol.json 2>&1 | tail -1 | cut -c1-1000; done
python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -k "config2" 2>&1 | tail -4
