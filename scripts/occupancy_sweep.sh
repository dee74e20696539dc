for pad in 0 4000 8000 13000 21000; do USIM_SMEM_PAD=$pad python bench.py --steps 150 --warmup 30 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pad $pad value', int(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4))"; done
