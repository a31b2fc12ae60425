mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_freeze.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r1c_pytest_freeze.log 2>&1; echo "freeze tests rc=$?"
tail -4 gpurun_out/r1c_pytest_freeze.log
timeout 120 python tools/bench_freeze.py > gpurun_out/r1c_freeze_bench.json 2> gpurun_out/r1c_freeze_bench.err; echo "bench_freeze rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1c_freeze_bench.json'))
for k,v in d.items():
    if isinstance(v,dict): print("%-50s %8.1f us %8.0f GB/s  %.2f"%(k,v['us'],v['GBs'],v['frac_hbm']))
PY
tail -3 gpurun_out/r1c_freeze_bench.err
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_freeze.py > gpurun_out/r1c_pytest_rest.log 2>&1; echo "rest tests rc=$?"
tail -4 gpurun_out/r1c_pytest_rest.log
timeout 200 python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; echo "bench rc=$?"; cat gpurun_out/r1c_bench.json
for w in decode_m1 decode_m8 decode_m32 int8_m4096 llama3_8b_decode_b1 llama3_8b_decode_b8; do
  timeout 150 python bench.py --workload $w > gpurun_out/r1c_bench_$w.json 2> gpurun_out/r1c_bench_$w.err; echo "bench $w rc=$?"; cat gpurun_out/r1c_bench_$w.json | cut -c1-900
done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"quantize_qbits_max|quantize_symmetric|quantize_qbytes_absmax" -c 3 -f -o gpurun_out/r1c_freeze python tools/bench_freeze.py --ncu > gpurun_out/r1c_ncu.log 2>&1; echo "ncu rc=$?"
