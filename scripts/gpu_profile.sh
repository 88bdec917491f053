#!/bin/bash
# ncu evidence for one workload: launch list (one step) + --set full capture of the main kernels of one step.
# Usage: bash scripts/gpu_profile.sh <tag> <workload>
TAG=$1; W=${2:-cfg2}
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
KRE="split_count|split_place|upsert_regions|emit_scan|emit_write_kernel|parse_lines_kernel|heads_lookup|heads_group_kernel"
timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-nohint > gpurun_out/bench_${TAG}_$W.json 2> gpurun_out/bench_${TAG}_$W.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${TAG}_$W.csv \
    python bench.py --workload $W --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-nohint > gpurun_out/ncu_launches_${TAG}_$W.log 2>&1; echo "ncu launches rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_${TAG}_$W.csv 4 > gpurun_out/launch_summary_${TAG}_$W.txt 2>&1; cat gpurun_out/launch_summary_${TAG}_$W.txt
# one launch of each main kernel per step (hint path, one chunk): skip the 3 warm-up steps
N=$(grep -cE "$KRE" gpurun_out/launches_${TAG}_$W.csv); PER=$((N / 4)); echo "matched launches: $N ($PER per step)"
timeout 1500 ncu --set full --clock-control none -k regex:"$KRE" -s $((3 * PER)) -c $((PER > 7 ? 7 : PER)) -f -o gpurun_out/prof_${TAG}_$W \
    python bench.py --workload $W --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-nohint > gpurun_out/ncu_full_${TAG}_$W.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | grep $TAG
