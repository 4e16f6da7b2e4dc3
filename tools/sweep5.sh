#!/bin/bash
cd "$(dirname "$0")/.."
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'match', d['config']['counts_match_between_device_and_host_api'])
except Exception as e: print('$1', 'FAILED', e)"; }
run "default"
for v in abm_amro_b200/lib_variants/*/; do AMRO_B200_LIB=$v/libamro_b200.so run "$(basename $v)"; done
