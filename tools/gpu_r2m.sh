#!/bin/bash
# round 2, GPU run M (1 GPU): the reference's OWN hot-path tests against the sm_100a kernels (bound into the real
# optimum.quanto classes), and the same-box comparison against the kernels the reference dispatches to on a B200
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python tools/run_reference_tests.py > gpurun_out/r2m_reference_tests.log 2>&1; echo "reference tests rc=$?"
tail -n 15 gpurun_out/r2m_reference_tests.log
timeout 600 python tools/compare_reference_kernels.py > gpurun_out/r2m_compare_reference_kernels.log 2>&1; echo "compare rc=$?"
tail -n 40 gpurun_out/r2m_compare_reference_kernels.log
ls gpurun_out/*.json 2>/dev/null | head
