#!/bin/bash
# Batched-kernel iteration: variant parity (debug listing + pytest), timing per variant, ncu --set full of one variant.
# Usage: bash tools/run_gpu_batched.sh <tag> [ncu_cfg]
TAG=${1:-b}
NCU_CFG=${2:-}
mkdir -p gpurun_out
timeout 300 python tools/dbg_batched.py > gpurun_out/dbg_batched_${TAG}.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -k "batched" > gpurun_out/pytest_batched_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_batched_${TAG}.log
timeout 300 python tools/gpu_probe.py batched > gpurun_out/probe_batched_${TAG}.log 2>&1
if [ -n "$NCU_CFG" ]; then
  BATCHED_CFG=$NCU_CFG timeout 600 ncu --set full --clock-control none --import-source on -k regex:batched_lu32 \
      -o gpurun_out/prof_batched_${TAG} -f python tools/gpu_probe.py batchedone > gpurun_out/ncu_batched_${TAG}.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_batched_${TAG}.ncu-rep > gpurun_out/ncu_batched_${TAG}_summary.txt 2>&1
fi
cat gpurun_out/dbg_batched_${TAG}.log
tail -5 gpurun_out/pytest_batched_${TAG}.log
cat gpurun_out/probe_batched_${TAG}.log
cat gpurun_out/ncu_batched_${TAG}_summary.txt 2>/dev/null
