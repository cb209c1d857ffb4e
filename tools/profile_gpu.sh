#!/usr/bin/env bash
# Run under gpurun: ncu launch list of the default bench command + one full capture of the dominant kernel.
# Usage: tools/profile_gpu.sh <tag> [extra bench args]
set -u
TAG=${1:-r01}; shift || true
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --batch 4 --no-cpu "$@" > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:segment -s 4 -c 2 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
