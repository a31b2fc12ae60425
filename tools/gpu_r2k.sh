#!/bin/bash
# round 2, GPU run K (1 GPU): second-generation ring gemv: parity tests, route timings, bench extras
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_bench_shapes.py -q -m gpu -x -k "ring or small_m or gemv" > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r2k_pytest.log
timeout 120 python tools/trace_gemvr.py 1 > gpurun_out/r2k_trace_gemvr_m1.log 2>&1; cat gpurun_out/r2k_trace_gemvr_m1.log
timeout 300 python tools/decode_routes.py > gpurun_out/r2k_decode_routes.log 2>&1; echo "routes rc=$?"; cat gpurun_out/r2k_decode_routes.log
