#!/bin/bash
# two B200: which part of the exchange still costs the slab step 0.07 ms?  (results are wrong with anything skipped)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { timeout 200 $TR bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','finite') if k in d}) for d in map(json.loads, sys.stdin)]"; }
{
for m in 0 1 2 4 7 3; do echo "== skip mask $m (1 qint planes, 2 pack/unpack, 4 copies+flags)"; IMHD_TMP_SKIP=$m run; done
} > gpurun_out/r2c_multi5.log 2>&1
cat gpurun_out/r2c_multi5.log
