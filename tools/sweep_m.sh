#!/bin/bash
# cfg5: inducing-point sweep at N = 1e6 rows/output, Q = 3, [Gaussian, Bernoulli, Poisson] (phase timings, full step)
for m in 64 128 256 512 1024 2048; do
  timeout 600 python tools/tc_check.py time sweepM$m ${1:-1000000} 2>&1 | grep -E "tc full|rror" | cut -c1-220
done
