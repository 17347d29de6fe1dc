#!/bin/bash
# GPU session M (1 GPU): host trace (BANI_TRACE) of two resident steps -- where the time between the timed stages goes
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
BANI_TRACE=1 BENCH_DEBUG=1 timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-files > $O/m_bench_n1.json 2> $O/m_bench_n1.err
grep -n "step\[res\]" $O/m_bench_n1.err | tail -3
