#!/bin/bash
mkdir -p gpurun_out
for c in all x2 m300; do
  timeout 300 python tools/tc_check.py small $c 2>&1 | grep -v Warning | tail -1 | cut -c1-200
done
timeout 900 python tools/tc_check.py time cfg3 1000000 2>&1 | grep -E "TIME" | cut -c1-400
