#!/bin/bash
# usage: bash scripts/multi_gpu_round.sh <N> <tag> [ppo]   -- bench.py on N GPUs of one box (torchrun, NCCL), optionally the PPO driver
N=$1; tag=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
tail -c 400 gpurun_out/bench_${tag}_n$N.json
if [ "$3" = "ppo" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 -m rui_b200.rl --config examples/rl_config_smoke.yaml --num-envs 65536 --n-steps 32 --total-timesteps 4.2e7 > gpurun_out/ppo_${tag}_n$N.log 2>&1
  tail -4 gpurun_out/ppo_${tag}_n$N.log
fi
