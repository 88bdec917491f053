#!/bin/bash
# bench the k=55 / k=91 single-GPU twins (KW=2, KW=3 kernels); optional ncu capture of the KW=2 extract kernel
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
for w in cfg4s cfg3s cfg5s; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${1:-x}_$w.json 2>> gpurun_out/workloads.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_${1:-x}_$w.json'))
print('$w', 'kmers/s=%.3g'%d['value'], 'ms=%.1f'%d['ms_per_step'], d['phase_ms_per_step'], 'e2e=%.3g'%d['e2e']['value'], 'frac=%.3f'%d['roofline']['frac'], d['config']['distinct_kmers'], d['table'])
"
done
tail -3 gpurun_out/workloads.err
