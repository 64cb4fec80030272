#!/bin/bash
# 1 -> N scaling run the way the driver launches it
set -u
mkdir -p gpurun_out
TAG=${1:-s}
NG=$(nvidia-smi -L | wc -l)
echo "gpus=$NG"
for N in 1 2 4 8; do
  [ "$N" -gt "$NG" ] && break
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-spmv > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_${TAG}_n$N.json") if l.startswith("{")][-1])
    print("  value", round(d["value"],1), "ms/step", round(d["ms_per_step"],1), "iters", d["detail"]["admm_iters_per_step"], "kcg", round(d["detail"]["pcg_iters_per_admm_iter"],2), "e2e", round(d["e2e"]["value"],1), "status", d["detail"]["status"])
    print("  phases", d["detail"].get("phase_ms_per_step") or d["detail"].get("phase_ms_per_step_rank0"))
except Exception as e:
    print("  parse failed", e)
PY
  tail -2 gpurun_out/scale_${TAG}_n$N.err | cut -c1-300
done
