#!/bin/bash
# GPU session E (8 GPUs): N = 8 and N = 4 bench lines
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
BENCH_DEBUG=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 2 > $O/e_bench_n8.json 2> $O/e_bench_n8.err
BENCH_DEBUG=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 2 > $O/e_bench_n4.json 2> $O/e_bench_n4.err
tail -3 $O/e_bench_n8.err
