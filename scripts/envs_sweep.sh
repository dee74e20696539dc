for n in 1184 2368 3552 4096 4736 8192 16384 32768; do python bench.py --steps 100 --warmup 20 --no-cpu --envs $n 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('envs $n value', int(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'us/env-wave', round(d['roofline']['kernel_ms']*1e3/($n/1184.0),2))"; done
