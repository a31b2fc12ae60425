#!/bin/bash
# round 2, GPU run T (1 GPU): sanity of the final tree (after host-only changes): decode tests, smoke, headline bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_cabi.py tests/test_gpu_freeze.py -q -m gpu -x -k "ring2 or absmax or affine" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r2t_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 200 python bench.py --no-extras --steps 20 --warmup 5 2> gpurun_out/r2t_bench.err | cut -c1-330
