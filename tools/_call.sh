# usage: bash tools/_call.sh <out-name> variant...   (config 2 only, 5 reps; development)
out=gpurun_out/$1; shift
mkdir -p gpurun_out; : > $out
for v in "$@"; do
  if [ "$v" = default ]; then python tools/time_config.py 30 10000 365 1000 5 2>&1 | sed "s/^/default: /" >> $out
  else AMRO_B200_LIB=abm_amro_b200/lib_variants/$v/libamro_b200.so python tools/time_config.py 30 10000 365 1000 5 2>&1 | sed "s/^/$v: /" >> $out; fi
done
cat $out
