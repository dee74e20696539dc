#!/bin/bash
# closed-loop statistics (reference `tracking` policy) for the shipped probe and for a fitted one; PPO bench; solver tolerance study
tag=${1:-r02d}
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/gputests_${tag}.log
python scripts/closed_loop_probe.py --envs 4096 --steps 3000 --json gpurun_out/closed_loop_${tag}_shipped.json > gpurun_out/closed_loop_${tag}_shipped.log 2>&1
python scripts/closed_loop_probe.py --envs 4096 --steps 3000 --probe-json gpurun_out/probe_fit_snapshot1.json --json gpurun_out/closed_loop_${tag}_fit1.json > gpurun_out/closed_loop_${tag}_fit1.log 2>&1
python bench.py --workload ppo --steps 5 --warmup 3 > gpurun_out/bench_${tag}_ppo.json 2> gpurun_out/bench_${tag}_ppo.err
for tol in 3e-5 1e-4; do python bench.py --steps 100 --warmup 10 --no-cpu --tol $tol 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tol $tol value', int(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'iters', round(d['config']['mean_solver_iters'],2))"; done
cat gpurun_out/gputests_${tag}.log
tail -32 gpurun_out/closed_loop_${tag}_shipped.log
tail -32 gpurun_out/closed_loop_${tag}_fit1.log
cat gpurun_out/bench_${tag}_ppo.json | cut -c1-1500; tail -3 gpurun_out/bench_${tag}_ppo.err
