#!/bin/bash
cd "$(dirname "$0")/.."
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], 'match', d['config']['counts_match_between_device_and_host_api'])"; }
for mod in 6 3 7; do for c in 0 2000 5000 10000 20000 40000; do AMRO_STAGGER_MOD=$mod AMRO_STAGGER_CYCLES=$c run "mod=$mod stagger=$c"; done; done
