#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__block_size --clock-control none -k regex:"k_" -s 10 -c 12 python tools/profile_step.py --steps 4 --path 1 2>&1 | grep -E "^\s+(void |imhd::|k_)|duration|grid_size" > gpurun_out/r2b_launches.txt
cat gpurun_out/r2b_launches.txt
