#!/bin/bash
# GPU session E (2 GPUs): multi-GPU path (overlapped sketch exchange, merged peer sketches), per-phase host timing
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > $O/e_bench_n2.json 2> $O/e_bench_n2.err
tail -5 $O/e_bench_n2.err
