#!/bin/bash
# One GPU session: tests, smoke, bench, launch list, ncu captures. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "$DO_REF" ]; then timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json; fi
if [ -n "$DO_NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$NCU_KERNELS" -s 6 -c 4 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $BENCH_ARGS > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out
fi
