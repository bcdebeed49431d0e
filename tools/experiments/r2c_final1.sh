#!/bin/bash
# round-2 final evidence on one B200: GPU test suite, bench legs, ncu full of one path B step, launch list, sanitizer
mkdir -p gpurun_out
tag=r02c
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${tag}_pytest_gpu.log; cat gpurun_out/${tag}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1_1000steps.json 2> gpurun_out/${tag}_bench_n1.err; cut -c1-600 gpurun_out/${tag}_bench_n1_1000steps.json; tail -3 gpurun_out/${tag}_bench_n1.err
timeout 300 python bench.py --workload pathA --no-cpu-baseline > gpurun_out/${tag}_bench_pathA_n1.json 2> gpurun_out/${tag}_bench_pathA.err; cut -c1-300 gpurun_out/${tag}_bench_pathA_n1.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ -s 16 -c 6 -f -o gpurun_out/${tag}_step_pathB \
    python tools/profile_step.py --steps 4 --path 1 > gpurun_out/${tag}_step_pathB.log 2>&1; tail -2 gpurun_out/${tag}_step_pathB.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_steps5.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
SAN_TOOLS="memcheck racecheck" SAN_PATHS="1" bash tools/sanitize.sh > gpurun_out/${tag}_compute_sanitizer.txt 2>&1; cat gpurun_out/${tag}_compute_sanitizer.txt
