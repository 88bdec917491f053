#!/bin/bash
# Compare kernel build variants (variants/lib_*.so) and the L2 fetch-granularity knob on the bench workload.
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
OUT=gpurun_out/tune_${1:-x}.jsonl
: > $OUT
run() { # name lib gran
  GENOMIX_GB_LIB=$2 GENOMIX_GB_L2_GRAN=$3 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/tune.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$1','gran':'$3','value':d['value'],'ms':d['ms_per_step'],'phase':d['phase_ms_per_step']}))" | tee -a $OUT
}
run default "" 0
run default "" 32
run default "" 128
for f in variants/lib_*.so; do n=$(basename $f .so); run $n $PWD/$f 0; run $n $PWD/$f 32; done
tail -3 gpurun_out/tune.err
