#!/bin/bash
# round 2, GPU run I (1 GPU): after the tcgen05.st fix: full GPU suite, smoke, e2e probe, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/r2i_status.txt
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_status.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2i_status.txt
timeout 200 python tools/e2e_probe.py > gpurun_out/r2i_e2e_probe.log 2>&1
timeout 120 python tools/trace_decode2.py 32 > gpurun_out/r2i_trace_decode_m32.log 2>&1
timeout 120 python tools/trace_decode2.py 8 > gpurun_out/r2i_trace_decode_m8.log 2>&1
echo "e2e_probe rc=$?" >> gpurun_out/r2i_status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2i_status.txt
cat gpurun_out/r2i_status.txt; tail -n 8 gpurun_out/r2i_pytest.log; tail -n 2 gpurun_out/r2i_smoke.log; cat gpurun_out/r2i_e2e_probe.log; cat gpurun_out/r2i_trace_decode_m32.log; tail -n 5 gpurun_out/r2i_bench_default.err
