#!/bin/bash
# Development: run tools/sweep.py <workload> <variants...> with the default library and with every tuning variant under
# abm_amro_b200/lib_variants/ (tools/build_variant.sh):   tools/sweep_libs.sh cfg2 "AMRO_BITS_G=18"
cfg=$1; shift
python tools/sweep.py $cfg "$@" | sed "s/^/default: /"
for d in abm_amro_b200/lib_variants/*/; do
  v=$(basename $d)
  AMRO_B200_LIB=$d/libamro_b200.so python tools/sweep.py $cfg "$@" | sed "s/^/$v: /"
done
