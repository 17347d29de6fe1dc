#!/bin/bash
# GPU session I (1 GPU, last of the round): parity tests + the bench line of the final code
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
timeout 420 python -m pytest tests -m gpu -q --durations=4 > $O/i_pytest.log 2>&1; echo "pytest rc $?" >> $O/i_pytest.log
timeout 400 python bench.py --steps 3 --warmup 3 > $O/i_bench_n1.json 2> $O/i_bench_n1.err
tail -2 $O/i_pytest.log
