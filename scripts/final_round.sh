#!/bin/bash
# final records of a round on one GPU: GPU suite, smoke(), default bench line, driver-style window, reference arm, PPO line
tag=${1:-r02}
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3 > gpurun_out/gputests_${tag}.log; cat gpurun_out/gputests_${tag}.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}_driver_style.json 2>> gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>> gpurun_out/bench_${tag}.err
python bench.py --workload ppo --steps 5 --warmup 3 > gpurun_out/bench_${tag}_ppo.json 2>> gpurun_out/bench_${tag}.err
python - <<PY
import json
for k in ("", "_driver_style", "_ref", "_ppo"):
    d = json.loads(open("gpurun_out/bench_${tag}%s.json" % k).read().strip().splitlines()[-1])
    print(k or "default", int(d["value"]), "e2e", int(d["e2e"]["value"]), "ms_per_step", round(d["ms_per_step"], 4), (d.get("roofline") or {}).get("kernel_ms"), (d.get("ppo") or {}).get("value"))
PY
