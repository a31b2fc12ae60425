#!/bin/bash
# round 2: ncu captures of the shipped kernels (decode ring gemv, tcgen05 decode kernel, int8 pair kernel, headline)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --set full --clock-control none --import-source on --graph-profiling node"
timeout 300 $NCU -k regex:gemv_w4s -s 30 -c 2 -f -o gpurun_out/r2_ncu_decode_m1 python bench.py --workload decode_m1 --steps 12 --warmup 3 > gpurun_out/r2_ncu_decode_m1.log 2>&1
echo "ncu m1 rc=$?"
timeout 300 $NCU -k regex:gemm_w4_decode -s 30 -c 2 -f -o gpurun_out/r2_ncu_decode_m32 python bench.py --workload decode_m32 --steps 12 --warmup 3 > gpurun_out/r2_ncu_decode_m32.log 2>&1
echo "ncu m32 rc=$?"
timeout 300 $NCU -k regex:gemm_tc2 -s 4 -c 2 -f -o gpurun_out/r2_ncu_int8 python bench.py --workload int8_m4096 --steps 6 --warmup 3 > gpurun_out/r2_ncu_int8.log 2>&1
echo "ncu int8 rc=$?"
timeout 300 $NCU -k regex:gemm_ -s 4 -c 2 -f -o gpurun_out/r2_ncu_int4_m4096 python bench.py --no-extras --steps 6 --warmup 3 > gpurun_out/r2_ncu_int4_m4096.log 2>&1
echo "ncu int4 rc=$?"
ls -la gpurun_out/*.ncu-rep
