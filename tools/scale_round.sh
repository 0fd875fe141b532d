#!/bin/bash
# 1 -> 8 GPU scaling of bench.py on one box (run under gpurun --gpus 8)
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_${n}gpu.json 2> gpurun_out/bench_r2_${n}gpu.err; echo "bench $n rc=$?"; cut -c1-200 gpurun_out/bench_r2_${n}gpu.json
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_r2_1gpu_same_box.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r2_1gpu_same_box.json
