#!/bin/bash
# GPU session H (1 GPU): final single-GPU artifacts -- parity tests, bench N = 1 (both arms), launch list, ncu --set full
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=6 > $O/h_pytest.log 2>&1; echo "pytest rc $?" >> $O/h_pytest.log
timeout 200 python tools/bench_stages.py 10 20 5000000 2 > $O/h_stages_200.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/h_bench_ref.json 2> $O/h_bench_ref.err
timeout 900 python bench.py --steps 3 --warmup 3 > $O/h_bench_n1.json 2> $O/h_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1000 -c 1400 --csv --log-file $O/h_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e-files > $O/h_bench_under_ncu.json 2> $O/h_bench_under_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"frag_l1|lookup_kernel|l2_events_kernel|l2_seq_kernel|l2_bounds_kernel|sketch_kernel|sort_unique|table_fill|zip_records|links_kernel" -c 45 -o $O/h_prof_map python tools/bench_stages.py 10 20 5000000 1 > $O/h_prof.log 2>&1
ls -la $O | tail -8
