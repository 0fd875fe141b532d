set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for v in "HMOGP_TC_FLUSH_ROWS=1024" "HMOGP_TC_FLUSH_ROWS=2048" "HMOGP_TC_FLUSH_ROWS=512" "HMOGP_TC_FLUSH_ROWS=1024 HMOGP_LIB=$PWD/hetmogp_b200/lib/var_foldlast.so"; do echo "== $v"; env $v python tools/tc_check.py scale cfg3 1000000 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc (full)|PARITY cfg3 N=1000000 tc vs" | cut -c1-400; done
