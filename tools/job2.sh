set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
for c in "cfg3 20000 tc" "cfg3 20000 fp32" "cfg4 20000 tc" "cfg2 20000 tc"; do timeout 600 python tools/oracle_check.py $c 2>&1 | tail -2; done > gpurun_out/oracle_check.log 2>&1; cat gpurun_out/oracle_check.log
