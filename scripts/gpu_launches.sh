#!/bin/bash
# ncu launch list (per-kernel device time) of one bench step.  Usage: gpu_launches.sh <tag> [bench args]
TAG=$1; shift
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
