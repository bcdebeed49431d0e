#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { timeout 200 $TR bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','finite') if k in d}) for d in map(json.loads, sys.stdin)]"; }
{
echo "== slab engine, one process per slab, direct exchange"; timeout 150 $TR tools/check_slab_engine.py 2>&1 | grep "bit-identical"
echo "== bench direct"; run
echo "== bench nccl"; IMHD_SLAB_EXCHANGE=nccl run
echo "== bench direct"; run
} > gpurun_out/r2c_multi6.log 2>&1
cat gpurun_out/r2c_multi6.log
