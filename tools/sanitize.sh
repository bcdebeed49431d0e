#!/bin/bash
# compute-sanitizer over a short fused run of both paths (the reference's only recorded check is a memcheck-clean run of
# its debug drivers, debug/build/debug.log).  Usage on the GPU box:  bash tools/sanitize.sh > gpurun_out/sanitizer.txt
set -u
cd "$(dirname "$0")/.."
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
for tool in ${SAN_TOOLS:-memcheck racecheck initcheck synccheck}; do
  for path in ${SAN_PATHS:-0 1}; do
    echo "=== $tool, path $path, 64x64x72, 2 fused steps (TMA kernel + remainder strip and z faces under it on the side stream + plane kernels)"
    IMHD_KERNEL_VARIANT=4 timeout 600 $CS --tool $tool --print-limit 5 python tools/profile_step.py --steps 2 --path $path --dims 64 64 72 2>&1 \
      | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|Barrier error|done" | head -12
  done
done
