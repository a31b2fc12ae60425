#!/bin/bash
# round 2, final 8-GPU run: bit-identity check of the column-parallel paths, then the driver's bench command (fused gather)
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/check_tp.py > gpurun_out/r2_check_tp_$N.log 2>&1; echo "check_tp rc=$?"
grep -v "^W0\|^\[W\|Warning\|detach\|return bool" gpurun_out/r2_check_tp_$N.log | tail -11
timeout 400 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_${N}_fused.json 2> gpurun_out/r2_scale_${N}_fused.err; echo "bench fused rc=$?"
tail -n 2 gpurun_out/r2_scale_${N}_fused.err | cut -c1-300
