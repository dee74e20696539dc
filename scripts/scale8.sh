#!/bin/bash
# eight-GPU record of the env step (config 4: 65 536 envs, 8192 per GPU) with the 1-GPU strong_ref, under torchrun
tag=${1:-r02}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 60 --warmup 10 > gpurun_out/bench_${tag}_n8.json 2> gpurun_out/bench_${tag}_n8.err
tail -c 700 gpurun_out/bench_${tag}_n8.json | head -c 400; echo; tail -3 gpurun_out/bench_${tag}_n8.err
