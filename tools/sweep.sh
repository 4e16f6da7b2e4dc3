#!/bin/bash
# Run the bench for the default build and every variant under abm_amro_b200/lib_variants (development tool).
cd "$(dirname "$0")/.."
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'match', d['config']['counts_match_between_device_and_host_api'])"; }
run default
for d in abm_amro_b200/lib_variants/*/; do AMRO_B200_LIB=$d/libamro_b200.so run $(basename $d); done
