for envs in 4096 8192; do for p in 0 1; do
USIM_PERSISTENT=$p python bench.py --steps 150 --warmup 20 --no-cpu --no-ppo --envs $envs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('persistent $p envs $envs value', int(d['value']), 'e2e', int(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'ms_per_step', round(d['ms_per_step'],4))"
done; done
USIM_PERSISTENT=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -k "sharding or determinism or auto_reset or back_to_back or soft_sweep" 2>&1 | tail -3
