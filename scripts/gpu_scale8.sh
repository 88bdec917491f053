#!/bin/bash
# The 8-GPU visit: multi-rank parity tests, the driver's bench line at N=8 (cfg2 weak + k=55 parity + cfg4 strong), and the two
# other multi-GPU configs of BASELINE.json as strong-scaled target workloads. Everything is kept under gpurun_out/.
TAG=${1:-r02}
N=${2:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpus_$TAG.txt
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
nvidia-smi nvlink -gt d > gpurun_out/nvlink_before_$TAG.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/pytest_mg_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_mg_$TAG.log
tail -3 gpurun_out/pytest_mg_$TAG.log
run() { # name args...
  name=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "$@" \
      > gpurun_out/bench_${TAG}_${name}_n$N.json 2> gpurun_out/bench_${TAG}_${name}_n$N.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_${TAG}_${name}_n$N.json") if l.startswith("{")][-1])
    print("$name", "value=%.3g" % d["value"], "ms=%.2f" % d["ms_per_step"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items() if v},
          "e2e=%.3g" % d["e2e"]["value"] if d.get("e2e") else "", "parity", (d.get("parity") or {}).get("ok"))
    t = d.get("target_workload")
    if t: print("   target", t["workload"][:40], "value=%.3g" % t["value"], "ms=%.1f" % t["ms_per_step"], {k: round(v, 1) for k, v in t["phase_ms_per_step"].items() if v})
except Exception as e:
    print("$name: no line", e)
PY
  tail -2 gpurun_out/bench_${TAG}_${name}_n$N.err
}
run main --steps 5 --warmup 3
nvidia-smi nvlink -gt d > gpurun_out/nvlink_after_main_$TAG.txt 2>&1
run cfg3 --steps 1 --warmup 3 --no-e2e --no-parity --target-workload cfg3 --target-steps 2
run cfg5 --steps 1 --warmup 3 --no-e2e --no-parity --target-workload cfg5 --target-steps 1 --target-stream
nvidia-smi nvlink -gt d > gpurun_out/nvlink_after_$TAG.txt 2>&1
du -sh gpurun_out
