#!/bin/bash
# Full ncu capture of kernels matching a regex during one bench step.
# Usage: gpu_ncu_kernel.sh <regex> <tag> [skip] [count] [extra bench args...]
RE=$1; TAG=$2; SKIP=${3:-3}; CNT=${4:-2}; shift 4 2>/dev/null || shift $#
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
