#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-b}
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "iters", d["config"]["admm_iters_per_step"], "kcg", d["config"]["pcg_iters_per_admm_iter"], "frac", d["roofline"]["frac"])
print("phases", d["config"].get("phase_ms_per_step"))
if "spmv" in d: print({k:(round(v["ms_flushed"]*1e3,1), round(v["gbs_flushed"])) for k,v in d["spmv"]["variants"].items()})
PY
tail -3 gpurun_out/bench_$TAG.err
