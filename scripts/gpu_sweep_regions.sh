#!/bin/bash
# upsert time vs number of table regions (cfg2, one B200)
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
OUT=gpurun_out/regions_${1:-x}.jsonl
: > $OUT
for r in 1 8 17 34 67 134 268 536 1024; do
  GENOMIX_GB_REGIONS=$r timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/regions.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'regions':$r,'ms':d['ms_per_step'],'phase':d['phase_ms_per_step']}))" | tee -a $OUT
done
