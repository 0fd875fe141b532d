set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python tools/tc_check.py scale cfg3 1000000 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc|PARITY" | cut -c1-400
python tools/oracle_check.py cfg3 20000 tc 2>&1 | tail -1 | cut -c1-1200
python tools/oracle_check.py cfg4 20000 tc 2>&1 | tail -1 | cut -c1-600
python tools/tc_check.py time cfg4 500000 2>&1 | grep -E "TIME" | cut -c1-300
