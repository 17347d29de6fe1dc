#!/bin/bash
# GPU session C: fixed grouping, ring 1024; staged vs direct l2_events; ncu of the new kernels; BASELINE configs 2/4/5 lines
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=12 > $O/c_pytest.log 2>&1; echo "pytest rc $?" >> $O/c_pytest.log
timeout 200 python tools/bench_stages.py 10 20 5000000 2 > $O/c_stages_200_staged.log 2>&1
timeout 200 python tools/bench_stages.py 10 20 5000000 2 l2_stage=0 > $O/c_stages_200_direct.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e-files > $O/c_bench_n1.json 2> $O/c_bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"frag_l1_kernel|lookup_kernel|l2_events_kernel|l2_seq_kernel|l2_bounds_kernel|sketch_kernel|sort_unique|table_fill|zip_records" -c 40 -o $O/c_prof_map python tools/bench_stages.py 10 20 5000000 1 > $O/c_prof.log 2>&1
timeout 900 python tools/bench_configs.py cfg2 cfg5 cfg4 100 > $O/c_configs.jsonl 2> $O/c_configs.err
ls -la $O | tail -10
