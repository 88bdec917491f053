#!/bin/bash
# bench sweep over one environment variable.  Usage: gpu_sweep_env.sh <tag> <VAR> <values...>
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
TAG=$1; VAR=$2; shift 2
OUT=gpurun_out/sweep_$TAG.jsonl
: > $OUT
for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${BENCH_ARGS:-} 2>>gpurun_out/sweep_$TAG.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'$VAR':'$v','ms':d['ms_per_step'],'phase':d['phase_ms_per_step'],'launches':d['gpu_launches']}))" | tee -a $OUT
done
tail -3 gpurun_out/sweep_$TAG.err
