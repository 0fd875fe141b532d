#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
for c in toy all m300 x2; do
  timeout 300 python tools/tc_check.py small $c 2>&1 | grep -v Warning | tail -1 | cut -c1-1500
done
timeout 300 python tools/tc_check.py small all fp32 2>&1 | grep -v Warning | tail -1 | cut -c1-1500
timeout 600 python tools/tc_check.py scale cfg3 20000 2>&1 | grep -E "PARITY|tc full"
timeout 900 python tools/tc_check.py scale cfg3 1000000 2>&1 | grep -v Warning | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gram_kernel|lik_rows_kernel" -c 6 -f -o gpurun_out/prof_tc python tools/tc_check.py time cfg3 200000 > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
