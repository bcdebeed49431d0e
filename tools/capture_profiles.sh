#!/bin/bash
# Round profile capture on one B200 (run under gpurun; outputs under gpurun_out/, summarised into profiles/ here):
#   tools/capture_profiles.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
# (a) one full time step, every kernel, ncu --set full: path B (6 launches per step after 4 set-up launches), path A
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 16 -c 6 -f -o gpurun_out/${tag}_step_pathB \
    python tools/profile_step.py --steps 4 --path 1 > gpurun_out/${tag}_step_pathB.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ -s 13 -c 3 -f -o gpurun_out/${tag}_step_pathA \
    python tools/profile_step.py --steps 4 --path 0 > gpurun_out/${tag}_step_pathA.log 2>&1
# (b) launch list of a short bench run (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_steps5.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
# (c) bench legs (never under a profiler)
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1_1000steps.json 2> gpurun_out/${tag}_bench_n1_1000steps.err
timeout 900 python bench.py --workload pathA --no-cpu-baseline > gpurun_out/${tag}_bench_pathA_n1.json 2> gpurun_out/${tag}_bench_pathA_n1.err
IMHD_C5_DIR=/tmp timeout 900 python bench.py --workload c5 --steps 200 --no-cpu-baseline > gpurun_out/${tag}_bench_c5_n1.json 2> gpurun_out/${tag}_bench_c5_n1.err
# (d) compute-sanitizer
bash tools/sanitize.sh > gpurun_out/${tag}_compute_sanitizer.txt 2>&1
tail -n 3 gpurun_out/${tag}_bench_n1_1000steps.err; cat gpurun_out/${tag}_bench_n1_1000steps.json | cut -c1-400
cat gpurun_out/${tag}_bench_pathA_n1.json | cut -c1-200; cat gpurun_out/${tag}_bench_c5_n1.json | cut -c1-300; tail -n 5 gpurun_out/${tag}_bench_c5_n1.err
grep -c "ERROR SUMMARY: 0" gpurun_out/${tag}_compute_sanitizer.txt; grep -E "ERROR SUMMARY|RACECHECK" gpurun_out/${tag}_compute_sanitizer.txt | sort | uniq -c
