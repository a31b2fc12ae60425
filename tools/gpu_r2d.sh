#!/bin/bash
# round 2, GPU run D: TMEM pair kernel first light, PDL decode, parity tests of the changed paths, timing table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/r2d_status.txt
timeout 120 python tools/check_w4p.py > gpurun_out/r2d_check_w4p.log 2>&1
echo "check_w4p rc=$?" >> gpurun_out/r2d_status.txt
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu -k "int4 or gather" > gpurun_out/r2d_pytest_bench_shapes.log 2>&1
echo "bench_shapes rc=$?" >> gpurun_out/r2d_status.txt
timeout 300 python -m pytest tests/test_gpu_cabi.py -x -q -m gpu -k "qbits" > gpurun_out/r2d_pytest_rest.log 2>&1
echo "rest rc=$?" >> gpurun_out/r2d_status.txt
timeout 250 python tools/gemv_modes.py > gpurun_out/r2d_gemv_modes.log 2>&1
echo "gemv_modes rc=$?" >> gpurun_out/r2d_status.txt
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2d_status.txt
cat gpurun_out/r2d_status.txt
tail -n 12 gpurun_out/r2d_check_w4p.log
for f in gpurun_out/r2d_pytest_bench_shapes.log gpurun_out/r2d_pytest_rest.log; do tail -n 4 $f; done
