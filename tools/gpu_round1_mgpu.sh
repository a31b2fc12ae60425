# usage: bash tools/gpu_round1_mgpu.sh <ngpus>   (under gpurun --gpus N)
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29541 tools/check_tp.py > gpurun_out/r1c_check_tp_$N.log 2>&1; echo "check_tp rc=$?"; grep -v "^W0\|^\[W\|Warning" gpurun_out/r1c_check_tp_$N.log | tail -12
timeout 200 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r1c_scale_${N}_fused.json 2> gpurun_out/r1c_scale_${N}_fused.err; echo "bench fused rc=$?"; cut -c1-1200 gpurun_out/r1c_scale_${N}_fused.json
timeout 200 $TR --master-port 29543 bench.py --gpus $N --steps 20 --warmup 3 --gather nccl > gpurun_out/r1c_scale_${N}_nccl.json 2> gpurun_out/r1c_scale_${N}_nccl.err; echo "bench nccl rc=$?"; cut -c1-600 gpurun_out/r1c_scale_${N}_nccl.json
timeout 100 $TR --master-port 29544 bench.py --gpus $N --steps 2 --warmup 1 --impl reference > gpurun_out/r1c_scale_${N}_ref.json 2> gpurun_out/r1c_scale_${N}_ref.err; echo "bench reference rc=$?"; cut -c1-300 gpurun_out/r1c_scale_${N}_ref.json
