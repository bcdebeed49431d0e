#!/bin/bash
# under-stream (strip + z faces co-resident with the marching kernel) vs one stream; ptxas register-usage levels
mkdir -p gpurun_out
{
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --tb=short -k "strip or ragged or plane_range or chunking or variants" 2>&1 | tail -5
echo "== under on (0) / off (8)"; timeout 600 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0,8,0,8 --check
for v in l0 l3 l5 l7 l10; do echo "== $v"; IMHD_B200_LIB=$PWD/tools/experiments/_build/libimhd_$v.so timeout 300 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0,8; done
echo "== launches"
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:"k_" -s 12 -c 7 python tools/profile_step.py --steps 4 --path 1 2>&1 | grep -E "^\s+(void |imhd::|k_)|duration" 
} > gpurun_out/r2c_ab1.log 2>&1
cat gpurun_out/r2c_ab1.log
