#!/bin/bash
# scaling bench on an N-GPU box: bash scripts/gpu_scale.sh <tag> <workload> <N...>
TAG=$1; W=$2; shift 2
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --steps 3 --warmup 3 --workload $W --no-cpu-baseline 2> gpurun_out/scale_${W}_n$N.err | grep "^{" > gpurun_out/bench_${TAG}_${W}_n$N.json
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --workload $W 2> gpurun_out/scale_${W}_n$N.err | grep "^{" > gpurun_out/bench_${TAG}_${W}_n$N.json
  fi
  python -c "
import json
d=json.load(open('gpurun_out/bench_${TAG}_${W}_n$N.json'))
print('$W N=$N', 'kmers/s=%.3g'%d['value'], 'ms=%.1f'%d['ms_per_step'], {k:round(v,2) for k,v in d['phase_ms_per_step'].items()}, 'e2e=%.3g'%d['e2e']['value'])
" || tail -5 gpurun_out/scale_${W}_n$N.err
done
