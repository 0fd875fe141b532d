set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python tools/tc_check.py time cfg3 1000000 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc (full)" | cut -c1-400
HMOGP_NO_GRAPH=1 python tools/tc_check.py time cfg3 1000000 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc (full)" | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_fwd_kernel|tc_gram2_kernel|tc_bwd_kernel" -c 3 -f -o gpurun_out/prof_r2a python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-variants --no-optimizer > gpurun_out/ncu_r2a.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r2a.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; cut -c1-3500 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
