#!/bin/bash
# multi-rank tests (virtual ranks on one GPU, real ranks when >1 GPU) + optional torchrun bench
set -u
mkdir -p gpurun_out
TAG=${1:-m}
NG=$(nvidia-smi -L | wc -l)
echo "gpus=$NG"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -p no:warnings > gpurun_out/pytest_multi_$TAG.log 2>&1; echo "pytest multi rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error|error|assert" gpurun_out/pytest_multi_$TAG.log | head -30
if [ "$NG" -gt 1 ]; then
  for N in 1 $NG; do
    if [ "$N" = "1" ]; then
      timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-spmv > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
    fi
    echo "bench N=$N rc=$?"; head -c 700 gpurun_out/bench_${TAG}_n$N.json; echo; tail -3 gpurun_out/bench_${TAG}_n$N.err
  done
fi
