#!/bin/bash
# Compare build variants (tune/lib_*.so) on several workloads.  Usage: bash scripts/gpu_tune3.sh <tag> <workload...>
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
TAG=${1:-x}; shift
OUT=gpurun_out/tune_$TAG.jsonl
: > $OUT
run() { # name lib workload
  GENOMIX_GB_LIB=$2 timeout 300 python bench.py --workload $3 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-nohint 2>>gpurun_out/tune_$TAG.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$1','workload':'$3','value':d['value'],'ms':d['ms_per_step'],'phase':d['phase_ms_per_step']}))" | tee -a $OUT
}
for w in "$@"; do
  run default "" $w
  for f in tune/lib_*.so; do n=$(basename $f .so); run $n $PWD/$f $w; done
done
tail -3 gpurun_out/tune_$TAG.err
