set -x
run() { echo "== $*"; env "$@" python tools/h_diag.py cfg3 200000 2>&1 | grep -E "BLOCKS|E relerr|rror" ; }
run X=1
run HMOGP_LIB=$PWD/hetmogp_b200/lib/var_nocentre.so
run HMOGP_TC_FLUSH_ROWS=1024
run HMOGP_TC_FLUSH_ROWS=2048
run HMOGP_TC_FLUSH_ROWS=256
run HMOGP_TC_FLUSH_ROWS=512 HMOGP_TC_FLUSH3_ROWS=4096
for v in "X=1" "HMOGP_LIB=$PWD/hetmogp_b200/lib/var_nocentre.so" "HMOGP_TC_FLUSH_ROWS=1024" "HMOGP_TC_FLUSH_ROWS=2048"; do echo "== $v"; env $v python tools/tc_check.py scale cfg3 1000000 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc full|PARITY cfg3 N=1000000 tc vs" | cut -c1-400; done
python tools/oracle_check.py cfg3 20000 tc 2>&1 | tail -1 | cut -c1-330
