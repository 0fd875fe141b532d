#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/tc_check.py time cfg3 1000000 2>&1 | grep -E "TIME" | cut -c1-400
timeout 900 ncu --set full --clock-control none -k regex:"tc_gram_kernel|tc_fwd_kernel" -s 2 -c 3 -f -o gpurun_out/prof_tc python tools/tc_check.py time cfg3 1000000 > gpurun_out/ncu_tc.log 2>&1; tail -1 gpurun_out/ncu_tc.log
