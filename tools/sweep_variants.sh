#!/bin/bash
# Development: time config 2 (and a many-particle configuration) with each tuning variant under abm_amro_b200/lib_variants/
cd "$(dirname "$0")/.."
for v in "$@"; do
  for cfg in "30 10000 365 1000" "20 5000 180 32768"; do
    if [ "$v" = default ]; then python tools/time_config.py $cfg | sed "s/^/default: /"
    else AMRO_B200_LIB=abm_amro_b200/lib_variants/$v/libamro_b200.so python tools/time_config.py $cfg | sed "s/^/$v: /"; fi
  done
done
