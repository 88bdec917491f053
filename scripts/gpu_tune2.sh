#!/bin/bash
# Compare build variants (tune/lib_*.so) on the bench workload.  Usage: bash scripts/gpu_tune2.sh <tag> [bench args]
mkdir -p gpurun_out
TAG=${1:-x}; shift
OUT=gpurun_out/tune_$TAG.jsonl
: > $OUT
run() { # name lib
  GENOMIX_GB_LIB=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline "${@:3}" 2>>gpurun_out/tune_$TAG.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$1','value':d['value'],'ms':d['ms_per_step'],'phase':d['phase_ms_per_step'],'table':d['table']}))" | tee -a $OUT
}
run default "" "$@"
for f in tune/lib_*.so; do n=$(basename $f .so); run $n $PWD/$f "$@"; done
tail -3 gpurun_out/tune_$TAG.err
