#!/bin/bash
# two B200: where does the slab step lose its 0.13 ms?  NCCL's send/recv kernels need SMs the interior launch holds.
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','finite') if k in d}) for d in map(json.loads, sys.stdin)]"; }
{
echo "== baseline"; run
echo "== 1 channel"; NCCL_MAX_NCHANNELS=1 NCCL_MIN_NCHANNELS=1 run
echo "== 2 channels"; NCCL_MAX_NCHANNELS=2 NCCL_MIN_NCHANNELS=2 run
echo "== copy engine"; NCCL_P2P_USE_CUDA_MEMCPY=1 run
echo "== 64 threads"; NCCL_NTHREADS=64 NCCL_MAX_NCHANNELS=2 run
echo "== no exchange (floor; wrong results)"; IMHD_TMP_NO_EXCHANGE=1 run
echo "== baseline, NCCL_DEBUG"; NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep -i "channel\|P2P\|nthreads" | head -20
} > gpurun_out/r2c_multi3.log 2>&1
cat gpurun_out/r2c_multi3.log
