#!/bin/bash
# One GPU-box visit: GPU test-suite, smoke, bench (both arms), ncu launch list and a full capture of the three N-sized kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 600 python bench.py --what ve --steps 10 > gpurun_out/bench_ve.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ve.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_fwd_kernel|tc_gram2_kernel|tc_bwd_kernel" -c 3 -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tc.log 2>&1; echo "ncu full rc=$?"
