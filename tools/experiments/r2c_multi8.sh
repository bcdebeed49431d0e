#!/bin/bash
# eight B200: direct exchange across eight processes (bit-identity), weak-scaling bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
{
echo "== slab engine, one process per slab, direct exchange"; timeout 150 $TR tools/check_slab_engine.py 2>&1 | grep "bit-identical\|rror"
} > gpurun_out/r2c_multi8.log 2>&1
timeout 240 $TR bench.py --gpus 8 --steps 300 --warmup 5 --no-cpu-baseline 2> gpurun_out/r02c_bench_n8.err | grep '^{' > gpurun_out/r02c_bench_n8.json
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_n8.json')); print({k:d[k] for k in ('value','ms_per_step','finite','n_gpus')}, d['e2e']['value'], d['clocks'])" >> gpurun_out/r2c_multi8.log 2>&1
cat gpurun_out/r2c_multi8.log; tail -3 gpurun_out/r02c_bench_n8.err
