set -x
for v in "HMOGP_TC_NPASS=3" "HMOGP_TC_NPASS=2" "HMOGP_TC_NPASS=1"; do echo "== $v"; env HMOGP_TC_FLUSH_ROWS=2048 $v timeout 300 python tools/tc_check.py time cfg3 1000000 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc (full)" | cut -c1-300; done
