#!/bin/bash
# One gpurun call that refreshes everything under profiles/: bench line, ncu launch list, full capture of the step launch.
# usage: bash scripts/profile_round.sh <tag>
tag=${1:-r01}
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>> gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 11 -c 1 -o gpurun_out/prof_solve_${tag} -f python bench.py --steps 3 --warmup 3 --no-cpu >> gpurun_out/bench_under_ncu_${tag}.log 2>&1
tail -c 600 gpurun_out/bench_${tag}.json
