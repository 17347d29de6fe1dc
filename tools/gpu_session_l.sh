#!/bin/bash
# GPU session L (1 GPU): BASELINE-scale parity tests on the final code (test_gpu.py ran in session K)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
timeout 170 python -m pytest tests/test_gpu_scale.py -q -m gpu --durations=4 > $O/l_pytest.log 2>&1; echo "pytest rc $?" >> $O/l_pytest.log
tail -n 8 $O/l_pytest.log
