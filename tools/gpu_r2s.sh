#!/bin/bash
# round 2, GPU run S (1 GPU): final validation of the last kernel changes: full GPU suite, smoke, default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 700 python -m pytest tests -q -m gpu > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2s_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2s_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench_default.json 2> gpurun_out/r2s_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2s_bench_default.json"))
print("head", round(d["value"],1), d["ms_per_step"], "frac", round(d["roofline"]["frac"],3), "e2e ms", d["e2e"]["ms_per_step"])
for k,v in d["extra"].items():
    print(k, round(v["value"],1), v["unit"], "ms", round(v["ms_per_step"],4), "frac", round(v["roofline"]["frac"],3)) if "error" not in v else print(k, v)
PY
