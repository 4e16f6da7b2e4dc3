#!/bin/bash
# The 8-GPU lines of profiles/ (run under `gpurun --gpus 8`):  tools/bench_multi8.sh <tag>
tag=${1:-run}; o=gpurun_out; mkdir -p $o
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --no-cpu-baseline"
$T --steps 20 --warmup 5 > $o/${tag}_bench_cfg2_30w_10kp_365d_1000s_weak_8gpu.json 2> $o/${tag}_8gpu.err
$T --workload cfg3_20w_5kp_180d_65536s --scaling strong --steps 5 --warmup 3 > $o/${tag}_bench_cfg3_20w_5kp_180d_65536s_strong_8gpu.json 2>> $o/${tag}_8gpu.err
$T --workload cfg5_5kw_10Mp_730d_32s --steps 3 --warmup 2 > $o/${tag}_bench_cfg5_5kw_10Mp_730d_32s_weak_8gpu.json 2>> $o/${tag}_8gpu.err
for f in $o/${tag}_bench_*8gpu.json; do tail -1 $f | cut -c1-200; done
