#!/bin/bash
# round 2, GPU run J (1 GPU): tensor-pipe probe for the decode kernel, reference-kernel comparison, reference tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 tools/mma_ts_probe > gpurun_out/r2j_mma_ts_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2j_mma_ts_probe.log
