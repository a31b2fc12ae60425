#!/bin/bash
# round 2, GPU run A: new parity tests, full suite, bench default + workloads, decode producer modes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -x -q -m gpu > gpurun_out/r2a_pytest_bench_shapes.log 2>&1
echo "bench_shapes rc=$?" >> gpurun_out/r2a_status.txt
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_bench_shapes.py > gpurun_out/r2a_pytest_rest.log 2>&1
echo "rest rc=$?" >> gpurun_out/r2a_status.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2a_status.txt
timeout 120 python tools/gemv_modes.py > gpurun_out/r2a_gemv_modes.log 2>&1
echo "gemv_modes rc=$?" >> gpurun_out/r2a_status.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2a_status.txt
cat gpurun_out/r2a_status.txt
for f in gpurun_out/r2a_pytest_bench_shapes.log gpurun_out/r2a_pytest_rest.log; do tail -n 3 $f; done
