#!/bin/bash
# round 2, GPU run H (1 GPU): state of the tree after the container restore: full GPU suite, smoke, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/r2h_status.txt
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_status.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2h_status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2h_status.txt
cat gpurun_out/r2h_status.txt; tail -n 5 gpurun_out/r2h_pytest.log; tail -n 2 gpurun_out/r2h_smoke.log; tail -n 5 gpurun_out/r2h_bench_default.err
