set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2b.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_pageable"]["value"]); print(d["variants"]); print(d["roofline"]["phase_ms_median"])
PY
