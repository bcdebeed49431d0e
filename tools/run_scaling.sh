#!/bin/bash
# Weak-scaling series of bench.py on one box: N = 1 (plain python) and N > 1 (torchrun), outputs under gpurun_out/.
#   tools/run_scaling.sh <tag> <steps> <N> [<N> ...]        e.g.  tools/run_scaling.sh r02 300 1 2 4 8
tag=$1; steps=$2; shift 2
mkdir -p gpurun_out
for n in "$@"; do
  out=gpurun_out/${tag}_bench_n${n}.json
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --steps $steps --warmup 5 ${BENCH_ARGS} > $out 2> gpurun_out/${tag}_bench_n${n}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps $steps --warmup 5 ${BENCH_ARGS} > $out 2> gpurun_out/${tag}_bench_n${n}.err
  fi
  echo "N=$n rc=$?"; python - "$out" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","finite","gpu_launches")}, "e2e", d["e2e"] and round(d["e2e"]["value"],2), "roofline", round(d["roofline"]["frac"],4), "clocks", d["clocks"])
except Exception as e:
    print("no JSON line:", e)
PY
done
