#!/bin/bash
# round 2, GPU run L (1 GPU): raw streaming ceiling of cp.async.bulk rings (tools/bulk_probe.cu), 29.36 MB over 148 CTAs
cd "$(dirname "$0")/../tools"
mkdir -p ../gpurun_out
{
for cfg in "4096 8 5 16 1 148 0 64" "4096 8 5 16 1 148 1 64" "4096 8 5 16 0 148 0 64" "32768 1 5 16 1 148 0 0" "16384 2 5 16 1 148 0 0" \
           "8192 4 5 16 1 148 0 0" "8192 4 5 16 1 148 1 0" "2048 16 5 16 1 148 0 64" "4096 8 3 16 1 148 0 64" "4096 4 10 16 1 148 0 64" \
           "4096 2 20 16 1 148 0 64" "4096 8 5 16 1 296 0 64" "4096 4 6 16 1 296 0 64"; do
  ./bulk_probe 29.36 $cfg
done
} > ../gpurun_out/r2l_bulk_probe.log 2>&1
cat ../gpurun_out/r2l_bulk_probe.log
