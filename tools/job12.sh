set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_${n}gpu.json 2> gpurun_out/bench_r2_${n}gpu.err; echo "bench $n rc=$?"; cut -c1-700 gpurun_out/bench_r2_${n}gpu.json; grep -o '"phase_ms_median": {[^}]*}' gpurun_out/bench_r2_${n}gpu.json; tail -3 gpurun_out/bench_r2_${n}gpu.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_r2_1gpu_same_box.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_1gpu_same_box.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 -m pytest tests/test_gpu_surface.py -x -q -m gpu -k "group" 2>&1 | tail -3
