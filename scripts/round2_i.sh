#!/bin/bash
python scripts/closed_loop_probe.py --envs 4096 --steps 3000 --json gpurun_out/closed_loop_r02i.json > gpurun_out/closed_loop_r02i.log 2>&1; tail -16 gpurun_out/closed_loop_r02i.log | cut -c1-330
python - <<PY
import sys, os, json
sys.path.insert(0, "scripts")
from closed_loop_probe import run
out = run(envs=1024, steps=1200)
json.dump(out, open("gpurun_out/closed_loop_r02i_small.json", "w"), indent=1)
import numpy as np
m, am, v, av = (np.array(out[k]) for k in ("obs_mean", "art_obs_mean", "obs_var", "art_obs_var"))
print("small run mean ratios", np.round(m[:6] / am[:6], 2), "var ratios", np.round(v[:12] / av[:12], 2), "ep_len", out["ep_len_mean"], "reward", out["step_reward"])
PY
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6
for k in 1 2; do python bench.py --steps 150 --warmup 30 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(int(d['value']), int(d['e2e']['value']), round(d['roofline']['kernel_ms'],4), round(d['config']['mean_solver_iters'],2), round(d['config']['mean_ncon'],1))"; done
