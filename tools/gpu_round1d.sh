mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_freeze.py -q -m gpu -p no:cacheprovider > gpurun_out/r1d_pytest_freeze.log 2>&1; echo "freeze tests rc=$?"
tail -4 gpurun_out/r1d_pytest_freeze.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/r1d_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r1d_smoke.log
timeout 120 python tools/bench_freeze.py > gpurun_out/r1d_freeze_bench.json 2> gpurun_out/r1d_freeze_bench.err; echo "bench_freeze rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1d_freeze_bench.json'))
for k,v in d.items():
    if isinstance(v,dict): print("%-50s %8.1f us %8.0f GB/s  %.2f"%(k,v['us'],v['GBs'],v['frac_hbm']))
PY
for w in decode_m1 decode_m8 decode_m32; do
  timeout 100 python bench.py --workload $w > gpurun_out/r1d_bench_$w.json 2> gpurun_out/r1d_bench_$w.err; echo "bench $w rc=$?"; cut -c1-1100 gpurun_out/r1d_bench_$w.json; tail -2 gpurun_out/r1d_bench_$w.err
done
timeout 100 python bench.py > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r1d_bench.json
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"quantize_qbits_max|quantize_symmetric|quantize_qbytes_absmax" -c 3 -f -o gpurun_out/r1d_freeze python tools/bench_freeze.py --ncu > gpurun_out/r1d_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python -m pytest tests -q -m gpu -p no:cacheprovider -x --deselect tests/test_gpu_freeze.py > gpurun_out/r1d_pytest_rest.log 2>&1; echo "rest tests rc=$?"
tail -3 gpurun_out/r1d_pytest_rest.log
