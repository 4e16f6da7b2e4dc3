#!/bin/bash
# Development: kernel time of config 2 / a 16k-particle sweep / a 1/5-scale config 4 for the default library and for each
# tuning variant under abm_amro_b200/lib_variants/ named on the command line (results go to profiles/sweeps.txt)
cd "$(dirname "$0")/.."
for v in default "$@"; do
  lib=abm_amro_b200/lib_variants/$v/libamro_b200.so; [ "$v" = default ] && lib=abm_amro_b200/lib/libamro_b200.so
  for w in cfg2 cfg3s cfg4s; do
    AMRO_B200_LIB=$lib python tools/sweep.py $w "" 2>&1 | tail -1 | sed "s/^/$v: /"
  done
done
