#!/bin/bash
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02m.json 2> gpurun_out/bench_r02m.err; tail -c 2500 gpurun_out/bench_r02m.json; tail -3 gpurun_out/bench_r02m.err
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-1500
python -c "import __graft_entry__ as g; g.smoke()"
