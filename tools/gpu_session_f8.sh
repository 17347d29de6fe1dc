#!/bin/bash
# GPU session F (8 GPUs): block partition at N = 8 / 4 / 2 and the round-robin deal at N = 8 for comparison
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
run() { n=$1; shift; BENCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 3 --warmup 2 "$@"; }
run 8 > $O/g_bench_n8.json 2> $O/g_bench_n8.err
run 4 > $O/g_bench_n4.json 2> $O/g_bench_n4.err
run 2 > $O/g_bench_n2.json 2> $O/g_bench_n2.err
run 8 --partition interleave > $O/g_bench_n8_interleave.json 2> $O/g_bench_n8_interleave.err
tail -2 $O/g_bench_n8.err
