#!/bin/bash
# The measured-and-rejected TMA staging of the trip inputs (kernels_bits.cu, -DAMRO_BITS_TMA): parity, timings, and one ncu capture
# on the HBM-streaming configuration (config 3).  Build first: tools/build_variant.sh tma -DAMRO_BITS_TMA
V=abm_amro_b200/lib_variants/tma/libamro_b200.so; o=gpurun_out
AMRO_B200_LIB=$V python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "philox or benchmarked or large_wards or teams_of" 2>&1 | tail -2
tools/sweep_r2.sh tma
B3="python bench.py --workload cfg3_20w_5kp_180d_65536s --steps 2 --warmup 1 --no-cpu-baseline"
AMRO_B200_LIB=$V $B3 > $o/r2_tma_plain3.json 2>$o/r2_tma.err && \
AMRO_B200_LIB=$V ncu --set full --clock-control none --import-source on -k regex:k_bits -s 1 -c 1 -f -o $o/r2_tma_k_bits_cfg3 $B3 > $o/r2_tma_ncu3.log 2>&1
python -c "import json; j=json.load(open('$o/r2_tma_plain3.json')); print('cfg3 tma', j['roofline']['kernel_ms'], j['roofline']['frac'])"
