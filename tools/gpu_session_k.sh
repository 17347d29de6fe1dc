#!/bin/bash
# GPU session K (1 GPU): dense identity tables restored in qsketch_map -- parity subset + N=1 bench line
cd "$(dirname "$0")/.." && mkdir -p gpurun_out && O=gpurun_out
BENCH_DEBUG=1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e-files > $O/k_bench_n1.json 2> $O/k_bench_n1.err
grep "step\[res\]" $O/k_bench_n1.err | head -4
timeout 150 python -m pytest tests/test_gpu.py -x -q -m gpu > $O/k_pytest.log 2>&1; tail -2 $O/k_pytest.log
python - <<'P'
import json
d=json.loads(open("gpurun_out/k_bench_n1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["value"], d["parity"], d["result_sha256"][:12], d.get("between_stages_ms"))
P
