#!/bin/bash
# round 2, multi-GPU run: usage  bash tools/gpu_r2_mgpu.sh <ngpus>   (under gpurun --gpus N)
# bit-identity of the column-parallel linear (NCCL and fused all-gather, chains, graphs), then the driver's bench command
# with the fused gather and with --gather nccl
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/check_tp.py > gpurun_out/r2_check_tp_$N.log 2>&1; echo "check_tp rc=$?"
grep -v "^W0\|^\[W\|Warning" gpurun_out/r2_check_tp_$N.log | tail -14
timeout 400 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_${N}_fused.json 2> gpurun_out/r2_scale_${N}_fused.err; echo "bench fused rc=$?"
tail -n 3 gpurun_out/r2_scale_${N}_fused.err | cut -c1-300
timeout 300 $TR --master-port 29543 bench.py --gpus $N --steps 20 --warmup 3 --gather nccl > gpurun_out/r2_scale_${N}_nccl.json 2> gpurun_out/r2_scale_${N}_nccl.err; echo "bench nccl rc=$?"
python - <<PY
import json
for tag in ("fused", "nccl"):
    try:
        d = json.load(open("gpurun_out/r2_scale_${N}_%s.json" % tag))
    except Exception as e:
        print(tag, "no line:", e); continue
    print(tag, "n_gpus", d["n_gpus"], "value %.1f %s" % (d["value"], d["unit"]), "ms %.4f" % d["ms_per_step"], "parity_ok", d.get("parity_ok"), d["config"]["parallelism"][:60])
    for k, v in d.get("extra", {}).items():
        if "error" in v: print("   ", k, v); continue
        print("   ", k, "%.1f %s" % (v["value"], v["unit"]), "ms %.4f" % v["ms_per_step"], "parity_ok", v.get("parity_ok"))
PY
