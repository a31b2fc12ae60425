mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_freeze.py -q -m gpu -p no:cacheprovider > gpurun_out/r1e_pytest_freeze.log 2>&1; echo "freeze tests rc=$?"
tail -4 gpurun_out/r1e_pytest_freeze.log
timeout 100 python tools/bench_freeze.py > gpurun_out/r1e_freeze_bench.json 2> gpurun_out/r1e_freeze_bench.err; echo "bench_freeze rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1e_freeze_bench.json'))
for k,v in d.items():
    if isinstance(v,dict): print("%-50s %8.1f us %8.0f GB/s  %.2f"%(k,v['us'],v['GBs'],v['frac_hbm']))
PY
timeout 50 python __graft_entry__.py smoke 2>&1 | tail -1
