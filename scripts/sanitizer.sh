#!/bin/bash
# compute-sanitizer memcheck + racecheck over scripts/sanitizer_driver.py (only our kernels: both attention backward kernels
# incl. query splits and phantom tiles, the 1-CTA and 2-CTA GEMMs with every fused epilogue, the memory-bound kernels).
# -> gpurun_out/sanitizer_$TAG.log
TAG=${1:-r2}
: > gpurun_out/sanitizer_$TAG.log
for TOOL in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $TOOL python scripts/sanitizer_driver.py" >> gpurun_out/sanitizer_$TAG.log
  timeout 420 compute-sanitizer --tool $TOOL --print-limit 10 python scripts/sanitizer_driver.py 2>&1 \
    | grep -v "Host Frame\|^=========\s*$" | head -60 >> gpurun_out/sanitizer_$TAG.log
done
cat gpurun_out/sanitizer_$TAG.log | cut -c1-240
