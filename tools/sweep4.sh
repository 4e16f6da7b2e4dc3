#!/bin/bash
cd "$(dirname "$0")/.."
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'match', d['config']['counts_match_between_device_and_host_api'])
except Exception as e: print('$1', 'FAILED', e)"; }
run "default wide768"
AMRO_B200_LIB=abm_amro_b200/lib_variants/w1024/libamro_b200.so run "wide1024"
AMRO_B200_LIB=abm_amro_b200/lib_variants/w512/libamro_b200.so run "wide512"
AMRO_FAST_SHAPE=t run "team128 auto G"
AMRO_FAST_SHAPE=t AMRO_FAST_G=5 run "team128 G=5"
