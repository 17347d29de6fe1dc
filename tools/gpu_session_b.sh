#!/bin/bash
# GPU session B (round 2): new kernels (lookup table, frag_l1, staged l2_events, warp sort/unique, packed ingest): parity + timings
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=30 > $O/b_pytest.log 2>&1; echo "pytest rc $?" >> $O/b_pytest.log
timeout 300 python tools/bench_stages.py 10 20 5000000 2 > $O/b_stages_200.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 2 > $O/b_bench_n1.json 2> $O/b_bench_n1.err
ls -la $O | tail -8
