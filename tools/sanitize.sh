#!/bin/bash
# compute-sanitizer over the smoke problem (all arithmetic modes; the CTA-pair tcgen05 kernels and the SIMT kernels).
# Each tool in its own process with its own timeout; summaries under gpurun_out/sanitizer_<tool>.log.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke |Error|hazard" gpurun_out/sanitizer_$tool.log | head -20
done
