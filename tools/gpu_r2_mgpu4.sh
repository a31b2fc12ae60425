#!/bin/bash
# round 2: the driver's bench command at 4 GPUs (fused gather)
cd "$(dirname "$0")/.."
N=4
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_${N}_fused.json 2> gpurun_out/r2_scale_${N}_fused.err; echo "bench fused rc=$?"
tail -n 2 gpurun_out/r2_scale_${N}_fused.err | cut -c1-300
