#!/bin/bash
# ncu --set full of one batched variant (both types), raw page + source page kept
CFG=${1:-256}
TAG=${2:-r2d}
mkdir -p gpurun_out
BATCHED_CFG=$CFG timeout 600 ncu --set full --clock-control none --import-source on -k regex:batched_lu32 \
    -o gpurun_out/prof_batched_${TAG} -f python tools/gpu_probe.py batchedone > gpurun_out/ncu_batched_${TAG}.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_batched_${TAG}.ncu-rep > gpurun_out/ncu_batched_${TAG}_summary.txt 2>&1
ncu -i gpurun_out/prof_batched_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_batched_${TAG}_raw.csv 2>/dev/null
cat gpurun_out/ncu_batched_${TAG}_summary.txt
