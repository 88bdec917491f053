#!/bin/bash
# One GPU-box visit (round 2): parity tests, bench line, ncu launch list, full captures of the top kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_r02.sh <tag> [tests-expr] [kernel-regex]
TAG=${1:-r02}
TESTS=${2:-tests}
KREGEX=${3:-"split_place|upsert_regions"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
timeout 1500 python -m pytest $TESTS -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launch_summary_$TAG.txt 2>&1; cat gpurun_out/launch_summary_$TAG.txt
if [ "${SKIP_FULL:-0}" != "1" ]; then
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-4} -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
ls -la gpurun_out/ | tail -20
fi
