#!/usr/bin/env bash
# Run under gpurun (round 2, fused colour path): launch list of the default bench command with the colour handling fused into the segment
# kernels and with the separate kernels (ACB200_FUSE=0), then one full steady-state capture (no cache control) of the fused segment kernels.
set -u
TAG=${1:-r02b}; shift || true
mkdir -p gpurun_out
for f in 1 0; do
ACB200_FUSE=$f ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}_fuse$f.csv \
    python bench.py --steps 2 --warmup 3 --batch 4 --no-cpu --no-extra --no-yuv "$@" > gpurun_out/bench_under_ncu_${TAG}_fuse$f.log 2>&1
done
ncu --set full --clock-control none --cache-control none --import-source on -k regex:segment_tm -s 8 -c 2 -o gpurun_out/prof_${TAG}_steady -f \
    python bench.py --steps 1 --warmup 3 --batch 2 --no-cpu --no-extra --no-yuv "$@" > gpurun_out/ncu_full_${TAG}_steady.log 2>&1
ls -la gpurun_out
