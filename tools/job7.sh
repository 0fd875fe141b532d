set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for v in "X=1" "HMOGP_TC_FLUSH_ROWS=1024" "HMOGP_TC_FLUSH_ROWS=4096" "HMOGP_LIB=$PWD/hetmogp_b200/lib/var_carry50.so" "HMOGP_LIB=$PWD/hetmogp_b200/lib/var_nocentre.so"; do echo "== $v"; env $v python tools/tc_check.py scale cfg3 1000000 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc (full)|PARITY cfg3 N=1000000 tc vs" | cut -c1-400; done
python tools/oracle_check.py cfg3 20000 tc 2>&1 | tail -1 | cut -c1-330
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; cut -c1-4500 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
