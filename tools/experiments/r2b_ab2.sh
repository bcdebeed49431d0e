#!/bin/bash
mkdir -p gpurun_out
{
echo "== variant_check 64"; timeout 300 python tools/experiments/variant_check.py 64
echo "== ab_bench default lib"; timeout 600 python tools/ab_bench.py --steps 100 --warmup 5 --paths B,A --variants 0,64,0,64 --check
echo "== ab_bench bare rcp"; IMHD_B200_LIB=$PWD/tools/experiments/_build/libimhd_nonewton.so timeout 600 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0,64 --check
echo "== parity suite, variant 64"; IMHD_KERNEL_VARIANT=64 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
} > gpurun_out/r2b_ab2.log 2>&1
cat gpurun_out/r2b_ab2.log
