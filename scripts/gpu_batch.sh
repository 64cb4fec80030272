#!/bin/bash
# Batched-path GPU run: smoke, batch parity tests, a standalone batch bench line.  Logs -> gpurun_out/
set -u
mkdir -p gpurun_out
TAG=${1:-batch}
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python -m pytest tests/test_gpu_batch.py -q -p no:warnings > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest batch rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_$TAG.log | head -40
timeout 600 python - > gpurun_out/batchbench_$TAG.json 2> gpurun_out/batchbench_$TAG.err <<'PY'
import sys, json, types
sys.path.insert(0, "osqp-python_b200"); sys.path.insert(0, ".")
from b200qp import batchbench
args = types.SimpleNamespace(steps=10, warmup=3)
print(json.dumps(batchbench.batch_line(args)))
PY
echo "batchbench rc=$?"; tail -c 2500 gpurun_out/batchbench_$TAG.json; tail -5 gpurun_out/batchbench_$TAG.err
