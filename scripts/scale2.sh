#!/bin/bash
# two-GPU records: env step (config 4, 65 536 envs sharded), PPO iteration (config 5), reference arm under torchrun
tag=${1:-r02}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 10 > gpurun_out/bench_${tag}_n2.json 2> gpurun_out/bench_${tag}_n2.err
$TR --master-port 29512 bench.py --gpus 2 --workload ppo --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ppo_n2.json 2> gpurun_out/bench_${tag}_ppo_n2.err
$TR --master-port 29513 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref_n2.json 2> gpurun_out/bench_${tag}_ref_n2.err
tail -c 600 gpurun_out/bench_${tag}_n2.json; echo; tail -c 400 gpurun_out/bench_${tag}_ppo_n2.json; echo; tail -c 300 gpurun_out/bench_${tag}_ref_n2.json
