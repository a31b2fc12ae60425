#!/bin/bash
# round 2, GPU run Q (1 GPU): the full GPU suite, smoke, the reference's own tests against the bound ops, Llama variants, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 TORCH_CUDA_ARCH_LIST=10.0
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2q_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2q_smoke.log
timeout 600 python tools/run_reference_tests.py > gpurun_out/r2q_reference_tests.log 2>&1; echo "reference tests rc=$?"; tail -n 2 gpurun_out/r2q_reference_tests.log
timeout 200 python tools/debug_ref_case2.py 2>&1 | tail -n 1
timeout 300 python tools/llama_variants.py 1 > gpurun_out/r2q_llama_variants.log 2>&1; cat gpurun_out/r2q_llama_variants.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench_default.json 2> gpurun_out/r2q_bench_default.err; echo "bench rc=$?"
