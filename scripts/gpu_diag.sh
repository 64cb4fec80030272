#!/bin/bash
set -u
mkdir -p gpurun_out
A="--steps 10 --warmup 3 --no-cpu-baseline --no-batch"
timeout 150 python bench.py $A --no-structured --no-spmv > gpurun_out/diag1.json 2> gpurun_out/diag1.err; echo "diag1 (solver only) rc=$?"; head -c 300 gpurun_out/diag1.json; echo
timeout 150 python bench.py $A --no-structured > gpurun_out/diag2.json 2> gpurun_out/diag2.err; echo "diag2 (+spmv) rc=$?"; tail -c 400 gpurun_out/diag2.json; echo
timeout 200 python bench.py $A --no-spmv > gpurun_out/diag3.json 2> gpurun_out/diag3.err; echo "diag3 (+structured) rc=$?"; tail -c 600 gpurun_out/diag3.json; echo
tail -3 gpurun_out/diag*.err
