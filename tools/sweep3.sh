#!/bin/bash
cd "$(dirname "$0")/.."
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'match', d['config']['counts_match_between_device_and_host_api'])
except Exception as e: print('$1', 'FAILED', e)"; }
for g in 7 1 2 4 5 6 8; do AMRO_FAST_CLUSTER=$g run "default(t128b6) G=$g"; done
for v in t256b3 t192b4 t512b1 t768b1; do for g in 1 2 3 4 7; do AMRO_B200_LIB=abm_amro_b200/lib_variants/$v/libamro_b200.so AMRO_FAST_CLUSTER=$g run "$v G=$g"; done; done
