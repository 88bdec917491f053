#!/bin/bash
# Multi-GPU visit: bash scripts/gpu_mg.sh <tag> <N> [extra bench args]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_$TAG.txt
nvidia-smi nvlink -gt d > gpurun_out/nvlink_before_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/pytest_mg_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_mg_$TAG.log
tail -5 gpurun_out/pytest_mg_$TAG.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench rc=$?"
nvidia-smi nvlink -gt d > gpurun_out/nvlink_after_$TAG.txt 2>&1
cat gpurun_out/bench_${TAG}_n$N.json | cut -c1-6000; tail -5 gpurun_out/bench_${TAG}_n$N.err
