#!/bin/bash
# First-contact GPU run: build check, smoke, sanitizer, gpu tests, short bench.  Logs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== build+smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== sanitizer (memcheck) on smoke"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
echo "== pytest gpu (small cases first)"
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_c4_full_size_properties --deselect tests/test_gpu_spmv.py::test_spmv_plan_linearity_at_scale > gpurun_out/pytest_gpu_small.log 2>&1; echo "pytest small rc=$?"; tail -15 gpurun_out/pytest_gpu_small.log
echo "== pytest gpu (full-size)"
timeout 900 python -m pytest tests/test_gpu_parity.py::test_c4_full_size_properties tests/test_gpu_spmv.py::test_spmv_plan_linearity_at_scale -q -x > gpurun_out/pytest_gpu_big.log 2>&1; echo "pytest big rc=$?"; tail -8 gpurun_out/pytest_gpu_big.log
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
