#!/bin/bash
# One GPU-box visit at the end of a round: GPU test-suite, smoke, bench (both arms, cfg2 / cfg4), ncu launch list, ncu full
# capture of the N-sized kernels, oracle check at the benchmark shape, Gram window sweep, sanitizer.  Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r2.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; echo "ref rc=$?"
for c in cfg2 cfg4; do timeout 600 python bench.py --config $c --no-cpu-baseline --steps 10 > gpurun_out/bench_r2_$c.json 2>> gpurun_out/bench_r2.err; cut -c1-200 gpurun_out/bench_r2_$c.json; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-variants > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_fwd_kernel|tc_gram2_kernel|tc_bwd_kernel|lik_rows_kernel" -c 8 -o gpurun_out/prof_r2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-parity --no-variants --no-optimizer > gpurun_out/ncu_tc.log 2>&1; echo "ncu full rc=$?"
(for c in "cfg3 20000 tc" "cfg3 20000 fp32" "cfg4 20000 tc" "cfg2 20000 tc"; do python tools/oracle_check.py $c 2>&1 | tail -2 | cut -c1-1500; done) > gpurun_out/oracle_check.log; cut -c1-300 gpurun_out/oracle_check.log
(for v in 512 1024 2048 4096; do echo "== HMOGP_TC_FLUSH_ROWS=$v"; env HMOGP_TC_FLUSH_ROWS=$v python tools/tc_check.py scale cfg3 1000000 2>&1 | grep -E "TIME cfg3 N=[0-9]* tc (full)|PARITY cfg3 N=1000000 tc vs" | cut -c1-400; done) > gpurun_out/gram_windows.txt; cut -c1-200 gpurun_out/gram_windows.txt
bash tools/sweep_m.sh > gpurun_out/sweep_m.txt 2>&1; cat gpurun_out/sweep_m.txt
SAN_TIMEOUT=600 bash tools/sanitize.sh
