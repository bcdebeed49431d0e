#!/bin/bash
mkdir -p gpurun_out
tag=r02c
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants or strip or plane_range or 100_steps" 2>&1 | tail -3
timeout 300 python tools/ab_bench.py --steps 100 --warmup 5 --paths B --variants 0,8
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1_1000steps.json 2> gpurun_out/${tag}_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_n1_1000steps.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['clocks'])"; tail -3 gpurun_out/${tag}_bench_n1.err
