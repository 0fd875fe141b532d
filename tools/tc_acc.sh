#!/bin/bash
# time (and accuracy with MODE=scale) of build variants / env settings at full size
#   tools/tc_acc.sh <variant|default>[,ENV=val,...] ...
for spec in "$@"; do
  IFS=',' read -ra parts <<< "$spec"
  v=${parts[0]}; envs=("${parts[@]:1}")
  if [ "$v" != default ]; then envs+=("HMOGP_LIB=$PWD/hetmogp_b200/lib/var_$v.so"); fi
  echo "== $spec"
  env "${envs[@]}" X=1 timeout 600 python tools/tc_check.py ${MODE:-time} cfg3 ${ROWS:-1000000} 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc|PARITY|rror" | cut -c1-420
done
