#!/bin/bash
mkdir -p gpurun_out
{
echo "== cur"; timeout 300 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0
for v in head head_args head_range l5 head; do echo "== $v"; IMHD_B200_LIB=$PWD/tools/experiments/_build/libimhd_$v.so timeout 300 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0; done
echo "== cur"; timeout 300 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0
} > gpurun_out/r2b_ab5.log 2>&1
cat gpurun_out/r2b_ab5.log
