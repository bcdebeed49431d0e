#!/bin/bash
# two B200: multi-GPU engine tests, then the slab step with the end launch beside (default) / in front of the interior launch
mkdir -p gpurun_out
{
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short 2>&1 | tail -5
for c in 1 0 1 0; do echo "== ends concurrent $c"; IMHD_SLAB_ENDS_CONCURRENT=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step','n_gpus','finite') if k in d}) for d in map(json.loads, sys.stdin)]"; done
} > gpurun_out/r2c_multi2.log 2>&1
cat gpurun_out/r2c_multi2.log
