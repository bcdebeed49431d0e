#!/bin/bash
mkdir -p gpurun_out
{
echo "== strip diff"; timeout 300 python tools/experiments/strip_diff.py
echo "== strip tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --tb=short -k "strip or ragged or plane_range or chunking" 2>&1 | tail -5
echo "== ab_bench"; timeout 600 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 256,0,256,0 --check
} > gpurun_out/r2b_ab3.log 2>&1
cat gpurun_out/r2b_ab3.log
