set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for v in "HMOGP_TC_NPASS=3" "HMOGP_TC_NPASS=1"; do echo "== $v"; env $v timeout 300 python tools/tc_check.py time cfg3 1000000 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc (full)" | cut -c1-300; done
