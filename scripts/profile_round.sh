#!/bin/bash
# One gpurun call that refreshes everything under profiles/: bench line, reference arm, ncu launch list, full capture of the step launch.
# usage: bash scripts/profile_round.sh <tag>     (short pre-roll under ncu: every launch is intercepted)
tag=${1:-r02}
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>> gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 260 -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 40 --warmup 10 --preroll 60 --no-cpu > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 150 -c 2 -o gpurun_out/prof_solve_${tag} -f python bench.py --steps 40 --warmup 10 --preroll 60 --no-cpu >> gpurun_out/bench_under_ncu_${tag}.log 2>&1
tail -c 800 gpurun_out/bench_${tag}.json
