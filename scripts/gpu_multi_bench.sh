#!/bin/bash
# bench.py the way the driver launches it at N > 1 (torchrun, one rank per GPU); N=$1, TAG=$2; extra env via ENVV="A=1 B=2"
set -u
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-m}
nvidia-smi -L | wc -l
env ${ENVV:-} timeout ${TMO:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
    bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:---no-batch} > gpurun_out/multi_${TAG}_n$N.json 2> gpurun_out/multi_${TAG}_n$N.err
echo "N=$N rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/multi_${TAG}_n$N.json") if l.startswith("{")][-1])
    print("  value %.1f e2e %.1f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["detail"])
except Exception as e:
    print("  parse failed", e)
PY
grep -v "^W\|Warning\|warn" gpurun_out/multi_${TAG}_n$N.err | tail -5 | cut -c1-300
