#!/bin/bash
# ncu captures of one bench step (run under gpurun, one GPU).  Usage: tools/profile.sh <tag>
# Writes gpurun_out/launches_<tag>.csv (launch list, gpu__time_duration) and
# gpurun_out/prof_{eval,scatter}_<tag>.ncu-rep (--set full, one launch each).
tag=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-configs > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 3 -c 1 -f -o gpurun_out/prof_eval_${tag} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-configs >> gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel -s 3 -c 1 -f -o gpurun_out/prof_scatter_${tag} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-configs >> gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out/
