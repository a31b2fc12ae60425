#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 TORCH_CUDA_ARCH_LIST=10.0
timeout 300 python tools/llama_variants.py 1 > gpurun_out/r2p_llama_variants.log 2>&1; echo "variants rc=$?"; cat gpurun_out/r2p_llama_variants.log
timeout 300 python tools/decode_routes.py 1 > gpurun_out/r2p_decode_routes.log 2>&1; cat gpurun_out/r2p_decode_routes.log
timeout 300 python tools/debug_ref_case2.py > gpurun_out/r2p_debug_ref_case2.log 2>&1; echo "rc=$?"; grep -v "^W0\|Warning\|Overriding\|operator:\|registered at\|dispatch key\|previous kernel\|new kernel\|self.m.impl" gpurun_out/r2p_debug_ref_case2.log | tail -12
