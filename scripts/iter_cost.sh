#!/bin/bash
# setup vs per-iteration cost of solve_kernel: bench with different CG iteration caps
for it in 1 2 4 8 40; do
  python bench.py --steps 60 --warmup 10 --no-cpu --iters $it 2>/dev/null | tail -1 > /tmp/b.json
  python - "$it" <<'PY'
import json, sys
d = json.load(open('/tmp/b.json'))
print("iters cap", sys.argv[1], "mean iters", round(d["config"]["mean_solver_iters"], 2), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "value", int(d["value"]))
PY
done
