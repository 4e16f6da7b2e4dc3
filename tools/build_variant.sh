#!/bin/bash
# Build a tuning variant of libamro_b200.so:  tools/build_variant.sh <name> <nvcc defines...>
# -> abm_amro_b200/lib_variants/<name>/libamro_b200.so   (select with AMRO_B200_LIB=<path>)
set -e
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
out=$here/abm_amro_b200/lib_variants/$name
mkdir -p "$out"
src=$here/abm_amro_b200/csrc
flags="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 --fmad=false $*"
for f in api kernels_fast kernels_bits kernels_general generate summary; do nvcc $flags -c $src/$f.cu -o $out/$f.o & done
g++ -O3 -std=c++17 -fPIC -ffp-contract=off -c $src/topology.cpp -o $out/topology.o &
wait
nvcc -shared -o $out/libamro_b200.so $out/api.o $out/kernels_fast.o $out/kernels_bits.o $out/kernels_general.o $out/generate.o $out/summary.o $out/topology.o
rm -f $out/*.o
echo "built $out/libamro_b200.so"
