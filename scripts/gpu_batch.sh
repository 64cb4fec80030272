#!/bin/bash
# Batched-path GPU run: batch parity tests, batch bench line in both inner-solver modes, ncu of the batch kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-batch}
timeout 900 python -m pytest tests/test_gpu_batch.py -q -p no:warnings > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest batch rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_$TAG.log | head -40
cat > /tmp/bb.py <<'PY'
import sys, json, types
sys.path.insert(0, "osqp-python_b200"); sys.path.insert(0, ".")
from b200qp import batchbench
args = types.SimpleNamespace(steps=int(sys.argv[1]), warmup=3)
print(json.dumps(batchbench.batch_line(args, with_cpu=len(sys.argv) > 2)))
PY
timeout 600 python /tmp/bb.py 10 cpu > gpurun_out/batchbench_${TAG}_dense.json 2> gpurun_out/batchbench_$TAG.err; echo "batchbench dense rc=$?"
if [ "${AB:-0}" = "1" ]; then
B200QP_BATCH_NT=1024 timeout 600 python /tmp/bb.py 10 > gpurun_out/batchbench_${TAG}_dense1024.json 2>> gpurun_out/batchbench_$TAG.err; echo "batchbench dense1024 rc=$?"
B200QP_BATCH_DENSE=0 timeout 600 python /tmp/bb.py 10 > gpurun_out/batchbench_${TAG}_pcg.json 2>> gpurun_out/batchbench_$TAG.err; echo "batchbench pcg rc=$?"
fi
python - <<PY
import json
import os
for k in [k for k in ("dense", "dense1024", "pcg") if os.path.exists("gpurun_out/batchbench_${TAG}_%s.json" % k)]:
    d = json.load(open("gpurun_out/batchbench_${TAG}_%s.json" % k))
    print(k, "QPs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 2), "it/qp", round(d["admm_iters_per_qp"], 1),
          "cg/it", round(d["pcg_iters_per_admm_iter"], 2), "grid", d["grid"], "solved", d["solved_fraction"], d.get("cpu_baseline", {}).get("value"))
PY
tail -3 gpurun_out/batchbench_$TAG.err
if [ "${PROFILE:-1}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:admm_batch_kernel -c 1 -f -o gpurun_out/prof_batch_$TAG \
      python /tmp/bb.py 1 > gpurun_out/ncu_batch_$TAG.log 2>&1; echo "ncu batch rc=$?"
  ncu -i gpurun_out/prof_batch_$TAG.ncu-rep --page details > gpurun_out/prof_batch_${TAG}_details.txt 2>&1
  ncu -i gpurun_out/prof_batch_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_batch_${TAG}_raw.csv 2>&1
  ls -la gpurun_out/*.ncu-rep 2>/dev/null
fi
