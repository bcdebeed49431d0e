#!/bin/bash
mkdir -p gpurun_out
{
echo "== cur l0"; timeout 300 python tools/ab_bench.py --steps 60 --warmup 5 --paths B --variants 0
for v in l5 l10 vb vb_l5 vb_l10; do echo "== $v"; IMHD_B200_LIB=$PWD/tools/experiments/_build/libimhd_$v.so timeout 300 python tools/ab_bench.py --steps 60 --warmup 5 --paths B --variants 0; done
} > gpurun_out/r2b_ab4.log 2>&1
cat gpurun_out/r2b_ab4.log
