#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fused_wstrip" -s 1 -c 1 -f -o gpurun_out/r2b_wstrip \
    python tools/profile_step.py --steps 3 --path 1 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
