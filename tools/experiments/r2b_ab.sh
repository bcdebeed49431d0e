#!/bin/bash
# Round-2b A/B run on one B200: carried-flux kernel (variant 48), co-resident strip (8), both (56), bare-rcp build.
mkdir -p gpurun_out
{
echo "== variant_check 48"; timeout 300 python tools/experiments/variant_check.py 48
echo "== ab_bench default lib"; timeout 600 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0,48,8,56,0,48 --check
echo "== ab_bench bare rcp"; IMHD_B200_LIB=$PWD/tools/experiments/_build/libimhd_nonewton.so timeout 600 python tools/ab_bench.py --steps 100 --warmup 5 --paths B,A --variants 0,48 --check
echo "== precision, bare rcp, variant 48"; IMHD_KERNEL_VARIANT=48 IMHD_B200_LIB=$PWD/tools/experiments/_build/libimhd_nonewton.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -s -k "c1_100_steps or second_order or bench_workload" 2>&1 | grep -E "normalised|passed|failed|rror" | tail -12
echo "== precision, default lib, variant 48"; IMHD_KERNEL_VARIANT=48 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s 2>&1 | grep -E "normalised|passed|failed|rror" | tail -12
} > gpurun_out/r2b_ab.log 2>&1
tail -40 gpurun_out/r2b_ab.log
echo "== full gpu suite, default"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2b_pytest_gpu.log
