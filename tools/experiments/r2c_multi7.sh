#!/bin/bash
# four B200: direct exchange with two distinct neighbours per slab (IPC), all-slabs-in-one-process tests, bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
run() { timeout 200 $TR bench.py --gpus 4 --steps 300 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','finite') if k in d}) for d in map(json.loads, sys.stdin)]"; }
{
echo "== slab engine, one process per slab, direct exchange"; timeout 150 $TR tools/check_slab_engine.py 2>&1 | grep "bit-identical\|rror"
echo "== tests (all slabs in one process)"; timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short 2>&1 | tail -5
echo "== bench direct, small blocks"; run
echo "== bench direct, big blocks"; IMHD_TMP_BIG_BLOCKS=1 run
echo "== bench nccl, small blocks"; IMHD_SLAB_EXCHANGE=nccl run
echo "== bench nccl, big blocks"; IMHD_TMP_BIG_BLOCKS=1 IMHD_SLAB_EXCHANGE=nccl run
} > gpurun_out/r2c_multi7.log 2>&1
cat gpurun_out/r2c_multi7.log
