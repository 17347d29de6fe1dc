#!/bin/bash
# GPU session J (1 GPU): per-step host phases of the final code (is the first timed resident step an outlier?)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
BENCH_DEBUG=1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e-files > $O/j_bench_n1.json 2> $O/j_bench_n1.err
grep "step\[res\]" $O/j_bench_n1.err | head -12
