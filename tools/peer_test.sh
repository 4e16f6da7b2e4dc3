#!/bin/bash
# development: the peer-store collective on 2 GPUs against NCCL (run under gpurun --gpus 2)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu-baseline --steps 20 --warmup 5"
P='import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "%.4e"%j["value"], round(j["ms_per_step"],3), j["roofline"]["kernel_ms_ranks"], "e2e %.4e"%j["e2e"]["value"], j["config"]["collective"][:60], j["gpu_launches"])'
$T 2>gpurun_out/peer.err | python -c "$P" peers; tail -3 gpurun_out/peer.err
AMRO_BENCH_NCCL=1 $T 2>/dev/null | python -c "$P" nccl
AMRO_BENCH_NO_COLLECTIVE=1 $T 2>/dev/null | python -c "$P" nocoll
