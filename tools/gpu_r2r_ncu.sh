#!/bin/bash
# round 2, GPU run R (1 GPU): ncu captures of the shipped kernels (second-generation ring gemv, tcgen05 decode kernel, int8
# pair kernel, headline TMEM pair kernel) condensed ON THE BOX into profiles-style summaries (the reports themselves exceed
# what gpurun copies back), and the launch list of the default bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --set full --clock-control none --graph-profiling node"
cap() {  # name, kernel regex, skip, bench args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 $NCU -k regex:$rx -s $skip -c 2 -f -o /tmp/$name python bench.py "$@" > gpurun_out/r2r_$name.log 2>&1
  echo "ncu $name rc=$?"
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/r2_${name}_ncu_summary.json "ncu --set full --clock-control none --graph-profiling node -k regex:$rx -s $skip -c 2 python bench.py $*"
}
cap decode_m1 gemv_w4r 30 --workload decode_m1 --steps 12 --warmup 3
cap decode_m8 gemv_w4r 30 --workload decode_m8 --steps 12 --warmup 3
cap decode_m32 gemm_w4_decode 30 --workload decode_m32 --steps 12 --warmup 3
cap int8 gemm_tc2 4 --workload int8_m4096 --steps 6 --warmup 3
cap int4_m4096 gemm_w4p 4 --no-extras --steps 6 --warmup 3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_default.csv python bench.py --no-extras --steps 2 --warmup 1 > gpurun_out/r2r_launches.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/ | tail -12
