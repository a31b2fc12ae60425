#!/bin/bash
# round 2, GPU run O (1 GPU): lite ring gemv (2 CTAs / SM), signed zero-points, the reference's tests and kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_bench_shapes.py tests/test_gpu_freeze.py -q -m gpu -x -k "ring or small_m or gemv or quantize_affine" > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/r2o_pytest.log
timeout 300 python tools/decode_routes.py 1 2 8 > gpurun_out/r2o_decode_routes.log 2>&1; echo "routes rc=$?"; cat gpurun_out/r2o_decode_routes.log
timeout 200 python bench.py --workload llama3_8b_decode_b1 --steps 10 --warmup 3 > gpurun_out/r2o_llama_b1.json 2> gpurun_out/r2o_llama_b1.err; echo "llama b1 rc=$?"; cut -c1-400 gpurun_out/r2o_llama_b1.json
timeout 200 python bench.py --workload llama3_8b_decode_b8 --steps 10 --warmup 3 > gpurun_out/r2o_llama_b8.json 2> gpurun_out/r2o_llama_b8.err; echo "llama b8 rc=$?"; cut -c1-400 gpurun_out/r2o_llama_b8.json
timeout 600 python tools/run_reference_tests.py > gpurun_out/r2o_reference_tests.log 2>&1; echo "reference tests rc=$?"
tail -n 3 gpurun_out/r2o_reference_tests.log
timeout 900 python tools/compare_reference_kernels.py > gpurun_out/r2o_compare_reference_kernels.log 2>&1; echo "compare rc=$?"
grep -v "^ours" gpurun_out/r2o_compare_reference_kernels.log | tail -n 40
