#!/bin/bash
# usage: bash scripts/variants_envs.sh <envs> <variant...>: bench every kernel variant built under build/ at a given env count
envs=$1; shift
for v in "$@"; do USIM_LIB=$PWD/build/libusim_$v.so python bench.py --steps 100 --warmup 20 --no-cpu --no-ppo --envs $envs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v envs $envs value', int(d['value']), 'e2e', int(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'iters', round(d['config']['mean_solver_iters'],2))"; done
