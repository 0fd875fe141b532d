#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
for c in all x2; do
  timeout 300 python tools/tc_check.py small $c 2>&1 | grep -v Warning | tail -1 | cut -c1-300
done
timeout 900 python tools/tc_check.py time cfg3 1000000 2>&1 | grep -v Warning | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gram_kernel" -c 1 -f -o gpurun_out/prof_tc python tools/tc_check.py time cfg3 200000 > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
