timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for c in "cfg3 1000000" "cfg2 100000" "cfg4 500000"; do timeout 300 python tools/tc_check.py time $c 2>&1 | grep -E "TIME cfg[0-9] N=[0-9]* tc (full)" | cut -c1-300; done
timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['phase_ms_median'])"
