#!/bin/bash
# 8-GPU visit: multi-rank parity (2 and 4 ranks), weak-scaling bench at N=8 for cfg2 (k=31) and cfg4s (k=55)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -5
for N in 8 4; do
for w in cfg2 cfg4s; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --workload $w 2> gpurun_out/scale_${w}_n$N.err | grep "^{" > gpurun_out/bench_${1:-x}_${w}_n$N.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_${1:-x}_${w}_n$N.json'))
print('$w N=$N', 'kmers/s=%.3g'%d['value'], 'ms=%.1f'%d['ms_per_step'], d['phase_ms_per_step'], 'e2e=%.3g'%d['e2e']['value'], d['config']['distinct_kmers'], d['table'])
" || tail -5 gpurun_out/scale_${w}_n$N.err
done; done
