#!/bin/bash
# usage: bash scripts/multi_gpu_round.sh <N> <tag>   (under gpurun --gpus N): config 4 scaling line (+ strong_ref) and config 5 PPO line
N=${1:-2}; tag=${2:-r02}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload ppo --steps 5 --warmup 3 > gpurun_out/bench_${tag}_ppo_n$N.json 2> gpurun_out/bench_${tag}_ppo_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref_n$N.json 2> gpurun_out/bench_${tag}_ref_n$N.err
for f in gpurun_out/bench_${tag}_n$N.json gpurun_out/bench_${tag}_ppo_n$N.json gpurun_out/bench_${tag}_ref_n$N.json; do echo "== $f"; tail -1 $f | cut -c1-1800; done
tail -3 gpurun_out/bench_${tag}_n$N.err gpurun_out/bench_${tag}_ppo_n$N.err | cut -c1-300
