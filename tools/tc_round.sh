#!/bin/bash
mkdir -p gpurun_out
for c in toy all x2 m300; do
  timeout 120 python tools/tc_check.py small $c 2>&1 | grep -v Warning | tail -1 | cut -c1-260
done
timeout 300 python tools/tc_check.py scale sweepM200 3000 2>&1 | grep -E "PARITY|rror" | cut -c1-400
timeout 600 python tools/tc_check.py scale cfg3 20000 2>&1 | grep -E "PARITY|rror" | cut -c1-400
timeout 600 python tools/tc_check.py scale cfg3 1000000 2>&1 | grep -E "TIME|PARITY|rror" | cut -c1-400
