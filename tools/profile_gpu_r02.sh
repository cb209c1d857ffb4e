#!/usr/bin/env bash
# Run under gpurun (round 2): launch list of the default bench command, one full capture of the luma segment kernels with the caches
# flushed between replays (cold) and one WITHOUT cache control (steady state: what the kernels see back to back).
set -u
TAG=${1:-r02}; shift || true
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --batch 4 --no-cpu --no-extra --no-yuv "$@" > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:segment_tm -s 8 -c 2 -o gpurun_out/prof_${TAG}_cold -f \
    python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu --no-extra --no-yuv "$@" > gpurun_out/ncu_full_${TAG}_cold.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:segment_tm -s 8 -c 2 -o gpurun_out/prof_${TAG}_steady -f \
    python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu --no-extra --no-yuv "$@" > gpurun_out/ncu_full_${TAG}_steady.log 2>&1
ncu --set full --clock-control none --cache-control none -k regex:"chroma_merge|rgb2yuv" -s 4 -c 2 -o gpurun_out/prof_${TAG}_pixel -f \
    python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu --no-extra --no-yuv "$@" > gpurun_out/ncu_full_${TAG}_pixel.log 2>&1
ls -la gpurun_out
