#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 TORCH_CUDA_ARCH_LIST=10.0
timeout 300 python tools/debug_ref_case.py > gpurun_out/r2n_debug_ref_case.log 2>&1; echo "rc=$?"; grep -v "^W0\|Warning" gpurun_out/r2n_debug_ref_case.log | tail -20
