set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
(python tools/h_diag.py cfg3 200000 > gpurun_out/hdiag_cfg3_200k.log 2>&1; tail -40 gpurun_out/hdiag_cfg3_200k.log)
(HMOGP_TC_FLUSH_ROWS=256 HMOGP_TC_FLUSH3_ROWS=4096 python tools/h_diag.py cfg3 200000 > gpurun_out/hdiag_cfg3_200k_w256.log 2>&1; grep -E "BLOCKS|E relerr" gpurun_out/hdiag_cfg3_200k_w256.log)
(python tools/h_diag.py cfg4 50000 > gpurun_out/hdiag_cfg4_50k.log 2>&1; grep -E "ELBO|BLOCKS|E relerr" gpurun_out/hdiag_cfg4_50k.log)
SAN_TIMEOUT=600 bash tools/sanitize.sh
