#!/bin/bash
# round 2, GPU run B: first light of the TMEM pair kernel, bench-shape parity tests, small-M routes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/r2b_status.txt
timeout 120 python tools/check_w4p.py > gpurun_out/r2b_check_w4p.log 2>&1
echo "check_w4p rc=$?" >> gpurun_out/r2b_status.txt
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu > gpurun_out/r2b_pytest_bench_shapes.log 2>&1
echo "bench_shapes rc=$?" >> gpurun_out/r2b_status.txt
timeout 300 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_qlinear.py -x -q -m gpu > gpurun_out/r2b_pytest_rest.log 2>&1
echo "rest rc=$?" >> gpurun_out/r2b_status.txt
timeout 200 python tools/gemv_modes.py > gpurun_out/r2b_gemv_modes.log 2>&1
echo "gemv_modes rc=$?" >> gpurun_out/r2b_status.txt
cat gpurun_out/r2b_status.txt
cat gpurun_out/r2b_check_w4p.log | tail -20
for f in gpurun_out/r2b_pytest_bench_shapes.log gpurun_out/r2b_pytest_rest.log; do tail -n 4 $f; done
