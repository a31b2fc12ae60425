#!/bin/bash
# round 2, GPU run F (1 GPU): timeline of the TMEM pair kernel, gather emulation through the new default route, ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/r2f_status.txt
timeout 120 python tools/trace_w4p.py > gpurun_out/r2f_trace_w4p.log 2>&1
echo "trace rc=$?" >> gpurun_out/r2f_status.txt
timeout 120 python tools/trace_w4p.py 4096 4096 14336 >> gpurun_out/r2f_trace_w4p.log 2>&1
timeout 400 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu -k "gather or int4_bench" > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_status.txt
bash tools/gpu_r2e_ncu.sh > gpurun_out/r2f_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2f_status.txt
cat gpurun_out/r2f_status.txt; cat gpurun_out/r2f_trace_w4p.log; tail -n 3 gpurun_out/r2f_pytest.log; tail -n 8 gpurun_out/r2f_ncu.log
