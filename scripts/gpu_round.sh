#!/bin/bash
# Regular GPU round: tests, bench, ncu launch list + full capture of the ADMM kernel.  Logs -> gpurun_out/
set -u
mkdir -p gpurun_out
TAG=${1:-r}
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q -p no:warnings > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu_$TAG.log | head -30
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
if [ "${PROFILE:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-spmv > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:admm_kernel -c 1 -f -o gpurun_out/prof_admm_$TAG \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-spmv > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
  ls -la gpurun_out/*.ncu-rep 2>/dev/null
fi
