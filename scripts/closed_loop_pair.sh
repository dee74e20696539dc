#!/bin/bash
# closed-loop statistics for one or more probe-parameter files: bash scripts/closed_loop_pair.sh <tag> <json> [<json> ...]
tag=$1; shift
for f in "$@"; do
  b=$(basename $f .json)
  python scripts/closed_loop_probe.py --envs 4096 --steps 3000 --probe-json $f --json gpurun_out/closed_loop_${tag}_${b}.json > gpurun_out/closed_loop_${tag}_${b}.log 2>&1
  echo "== $b"; tail -14 gpurun_out/closed_loop_${tag}_${b}.log | cut -c1-330
  python - <<PY
import json
d=json.load(open("gpurun_out/closed_loop_${tag}_${b}.json"))
names=["Fx","Fy","Fz","tq_x","tq_y","tq_z","vx","vy","vz","Fzm","dFz","velm","ex","ey","ez"]
print("mean ours:", " ".join(f"{n}={d['obs_mean'][i]:.3g}" for i,n in enumerate(names)))
print("mean ART :", " ".join(f"{n}={d['art_obs_mean'][i]:.3g}" for i,n in enumerate(names)))
print("var  ours:", " ".join(f"{n}={d['obs_var'][i]:.3g}" for i,n in enumerate(names)))
print("var  ART :", " ".join(f"{n}={d['art_obs_var'][i]:.3g}" for i,n in enumerate(names)))
PY
done
