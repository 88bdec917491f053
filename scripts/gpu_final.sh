#!/bin/bash
# What the driver runs at round end, in one visit: smoke, the GPU tests, both bench arms; plus one ncu capture.
TAG=${1:-final}
mkdir -p gpurun_out
export GX_BENCH_TEXT_CACHE=/tmp/gxtext
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "reference rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"emit_write_kernel" -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_emit_write \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-nohint > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.4g ms %.2f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if v})
print("no_hint", d["no_hint"]["ms_per_step"], d["no_hint"]["table"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
r = json.load(open("gpurun_out/bench_ref_$TAG.json")); print("reference arm %.4g" % r["value"])
PY
