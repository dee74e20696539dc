#!/bin/bash
# one gpurun call: GPU test suite, measured parity drift, two bench windows (phase independence), optional kernel variants
tag=${1:-r02a}; shift
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -40 > gpurun_out/gputests_${tag}.log
python scripts/parity_report.py --out gpurun_out/parity_drift_${tag}.json > gpurun_out/parity_${tag}.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_${tag}_short.json 2> gpurun_out/bench_${tag}.err
python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/bench_${tag}_long.json 2>> gpurun_out/bench_${tag}.err
cat gpurun_out/gputests_${tag}.log
tail -12 gpurun_out/parity_${tag}.log
python - <<PY
import json
for k in ("short","long"):
    try:
        d=json.loads(open("gpurun_out/bench_${tag}_%s.json"%k).read().strip().splitlines()[-1])
        print(k,'value',int(d['value']),'e2e',int(d['e2e']['value']),'ms_per_step',round(d['ms_per_step'],4),'kernel_ms',round(d['roofline']['kernel_ms'],4),'iters',round(d['config']['mean_solver_iters'],2),'ncon',round(d['config']['mean_ncon'],1))
    except Exception as e: print(k,'failed',e)
PY
tail -5 gpurun_out/bench_${tag}.err
[ $# -gt 0 ] && bash scripts/variants.sh "$@"
