#!/bin/bash
# round 2, GPU run G (1 GPU): re-validate the restructured staging loops (pair kernel + tcgen05 decode kernel), timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/r2g_status.txt
timeout 120 python tools/check_w4p.py > gpurun_out/r2g_check_w4p.log 2>&1
echo "check_w4p rc=$?" >> gpurun_out/r2g_status.txt
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_cabi.py -q -m gpu -k "int4 or gather or qbits" > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_status.txt
timeout 120 python tools/trace_w4p.py > gpurun_out/r2g_trace_w4p.log 2>&1
timeout 250 python tools/gemv_modes.py > gpurun_out/r2g_gemv_modes.log 2>&1
echo "gemv_modes rc=$?" >> gpurun_out/r2g_status.txt
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2g_status.txt
cat gpurun_out/r2g_status.txt; cat gpurun_out/r2g_check_w4p.log | tail -6; cat gpurun_out/r2g_trace_w4p.log | head -6; tail -n 3 gpurun_out/r2g_pytest.log; grep -A5 "large M" gpurun_out/r2g_gemv_modes.log
