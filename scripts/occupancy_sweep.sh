#!/bin/bash
# throughput of the env step against resident CTAs per SM (USIM_SMEM_PAD trades them away): how much would MORE of them buy?
for envs in 32768 4096; do
for pad in 0 1024 5200 10700; do
USIM_SMEM_PAD=$pad python bench.py --steps 100 --warmup 20 --no-cpu --no-ppo --envs $envs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('envs $envs pad $pad value', int(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'ms_per_step', round(d['ms_per_step'],4))"
done; done
