#!/bin/bash
# Full ncu capture of kernels matching a regex during one bench step.  Usage: gpu_ncu_kernel.sh <regex> <tag> [skip] [count]
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-3} -c ${4:-2} -f -o gpurun_out/prof_$2 \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$2.log 2>&1; echo "ncu rc=$?"
