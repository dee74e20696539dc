#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -k "config3_soft or config2 or reset_matches or ur5e or wrench" 2>&1 | tail -3
run() { python bench.py --steps 150 --warmup 20 --no-cpu "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$V value', int(d['value']), 'ms_per_step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'step-kernel us', round(1000*(d['ms_per_step']-d['roofline']['kernel_ms']),1))"; }
V=rolled run; V=rolled run
V=unrolled USIM_LIB=$PWD/build/libusim_unrolled.so run; V=unrolled USIM_LIB=$PWD/build/libusim_unrolled.so run
