timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for m in 64 128 256; do timeout 600 python tools/tc_check.py time sweepM$m 1000000 2>&1 | grep -E "tc full|rror" | cut -c1-220; done
for c in "cfg3 1000000" "cfg2 100000"; do timeout 300 python tools/tc_check.py time $c 2>&1 | grep -E "TIME cfg[0-9] N=[0-9]* tc (full)" | cut -c1-300; done
