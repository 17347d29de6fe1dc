#!/bin/bash
# GPU session A (round 2): parity tests, bench (both arms), launch list, ncu --set full of the map kernels, sanitizer.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
nproc > $O/a_host.txt; cat /sys/fs/cgroup/cpu.max >> $O/a_host.txt 2>&1; lscpu | grep -i "model name\|^CPU(s)" >> $O/a_host.txt; df -h /tmp | tail -1 >> $O/a_host.txt; free -g | head -2 >> $O/a_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/a_pytest.log 2>&1; echo "pytest rc $?" >> $O/a_pytest.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/a_bench_ref.json 2> $O/a_bench_ref.err
timeout 900 python bench.py --steps 3 --warmup 3 > $O/a_bench_n1.json 2> $O/a_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/a_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e-files > $O/a_bench_under_ncu.json 2> $O/a_bench_under_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"frag_l1_kernel|lookup_kernel|l2_events_kernel|l2_seq_kernel|l2_bounds_kernel|sketch_kernel" -c 24 -o $O/a_prof_map python tools/bench_stages.py 10 20 5000000 1 > $O/a_prof.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/a_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $O/a_racecheck.log 2>&1
ls -la $O | tail -20
