#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fused_split|k_fused_strip" -s 2 -c 2 -f -o gpurun_out/r2b_split_strip \
    python tools/profile_step.py --steps 3 --path 1 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log
timeout 300 python tools/ab_bench.py --steps 100 --warmup 5 --paths B,A --variants 0,32 --check
