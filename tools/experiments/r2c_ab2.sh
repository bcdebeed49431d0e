#!/bin/bash
# ptxas register-usage levels around 5; strip chunking under the marching kernel (warps per SM worth of strip tasks)
mkdir -p gpurun_out
{
echo "== default build (l5), under on / off"; timeout 600 python tools/ab_bench.py --steps 100 --warmup 5 --paths B,A --variants 0,8 --check
for v in l4 l6; do echo "== $v"; IMHD_B200_LIB=$PWD/tools/experiments/_build/libimhd_$v.so timeout 300 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0; done
for w in 3 4 6 12 16; do echo "== wpsm $w"; IMHD_WSTRIP_WPSM=$w timeout 300 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0; done
} > gpurun_out/r2c_ab2.log 2>&1
cat gpurun_out/r2c_ab2.log
