mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r1b_gpu.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_freeze.py -q -m gpu -p no:cacheprovider > gpurun_out/r1b_pytest_freeze.log 2>&1; echo "freeze tests rc=$?" 
tail -5 gpurun_out/r1b_pytest_freeze.log
timeout 120 python tools/bench_freeze.py > gpurun_out/r1b_freeze_bench.json 2> gpurun_out/r1b_freeze_bench.err; echo "bench_freeze rc=$?"
cat gpurun_out/r1b_freeze_bench.json | head -80
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_freeze.py --durations=8 > gpurun_out/r1b_pytest_rest.log 2>&1; echo "rest tests rc=$?"
tail -15 gpurun_out/r1b_pytest_rest.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"quantize_qbits_max|quantize_symmetric|quantize_qbytes_absmax" -c 3 -f -o gpurun_out/r1b_freeze python tools/bench_freeze.py --ncu > gpurun_out/r1b_ncu.log 2>&1; echo "ncu rc=$?"
timeout 200 python bench.py > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; echo "bench rc=$?"; cat gpurun_out/r1b_bench.json
