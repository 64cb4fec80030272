#!/bin/bash
# Round-2 GPU round: smoke, gpu tests, bench (both arms), ncu launch list, ncu full capture of the hot kernel with
# the raw / details pages exported on the box (the .ncu-rep comes back as well).  Logs -> gpurun_out/
#   TAG=$1   STEPS: PYTEST=1 BENCH=1 REFARM=1 PROFILE=1 (set to 0 to skip)
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_$TAG.txt 2>&1; nproc >> gpurun_out/gpu_$TAG.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
if [ "${PYTEST:-1}" = "1" ]; then
  timeout ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -q -p no:warnings --durations=15 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
  grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_$TAG.log | head -30
fi
if [ "${BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["detail"])
    r = d["roofline"]; print("frac", r["frac"], "engine", r["engine_model"]["frac"], r["per_iteration_spmv"])
    if "spmv" in d: print({k: (round(v["ms_flushed"] * 1e3, 1), round(v["gbs_flushed"])) for k, v in d["spmv"].get("variants", {}).items()} or d["spmv"])
    if "structured" in d: print("structured", {k: v for k, v in d["structured"].items() if k in ("value", "roofline", "error")})
    if "batch" in d: print("batch", {k: v for k, v in d["batch"].items() if k in ("value", "e2e", "error")})
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
  tail -3 gpurun_out/bench_$TAG.err
fi
if [ "${REFARM:-1}" = "1" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref arm rc=$?"; head -c 600 gpurun_out/bench_ref_$TAG.json
fi
if [ "${PROFILE:-1}" = "1" ]; then
  B200QP_AUTOTUNE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-spmv --no-batch --no-structured > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
  bash scripts/ncu_traffic.sh $TAG
fi
