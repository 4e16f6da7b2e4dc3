#!/bin/bash
# N-GPU bench line the way the driver launches it:  tools/bench_multi.sh N <tag>
N=$1; tag=${2:-run}; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
tail -1 gpurun_out/${tag}_bench_${N}gpu.json | cut -c1-300
