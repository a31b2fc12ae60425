#!/bin/bash
# round 2, multi-GPU run: NG ranks (argument, default 2): parity of the fused gather, bench scaling lines
cd "$(dirname "$0")/.."
NG=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=r2c_n${NG}
rm -f gpurun_out/${TAG}_status.txt
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 bash -c "$(declare -f run); NG=$NG; run 29541 tools/check_tp.py" > gpurun_out/${TAG}_check_tp.log 2>&1
echo "check_tp rc=$?" >> gpurun_out/${TAG}_status.txt
timeout 400 bash -c "$(declare -f run); NG=$NG; run 29542 bench.py --gpus $NG --steps 20 --warmup 5" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_status.txt
timeout 300 bash -c "$(declare -f run); NG=$NG; run 29543 bench.py --gpus $NG --steps 20 --warmup 5 --gather nccl --no-extras" > gpurun_out/${TAG}_bench_nccl.json 2> gpurun_out/${TAG}_bench_nccl.err
echo "bench nccl rc=$?" >> gpurun_out/${TAG}_status.txt
cat gpurun_out/${TAG}_status.txt
tail -n 12 gpurun_out/${TAG}_check_tp.log
tail -c 1500 gpurun_out/${TAG}_bench.err
