#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list and one full capture of the dominant kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_ref_$TAG.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:extract_kernel -s 4 -c 2 -f -o gpurun_out/prof_extract_$TAG \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/
fi
