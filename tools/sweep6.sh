#!/bin/bash
cd "$(dirname "$0")/.."
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'match', d['config']['counts_match_between_device_and_host_api'])
except Exception as e: print('$1', 'FAILED', e)"; }
export AMRO_FAST_SHAPE=t
run "team128b6 auto"
for g in 5 6 7; do AMRO_FAST_G=$g run "team128b6 G=$g"; done
for v in tt192 tt256 tt96 tt128b5; do AMRO_B200_LIB=abm_amro_b200/lib_variants/$v/libamro_b200.so run "$v auto"; done
AMRO_B200_LIB=abm_amro_b200/lib_variants/tt192/libamro_b200.so AMRO_FAST_G=4 run "tt192 G=4"
AMRO_B200_LIB=abm_amro_b200/lib_variants/tt256/libamro_b200.so AMRO_FAST_G=3 run "tt256 G=3"
AMRO_B200_LIB=abm_amro_b200/lib_variants/tt96/libamro_b200.so AMRO_FAST_G=9 run "tt96 G=9"
