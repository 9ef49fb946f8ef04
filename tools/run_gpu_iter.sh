#!/bin/bash
# Iteration run: parity tests (no full-size), selected probe sections, optional ncu of big gemm.
# Usage: bash tools/run_gpu_iter.sh <tag> "<probe sections>" [ncu]
TAG=${1:-it}
SECS=${2:-"gemm panel getrf getrs"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -k "not fullsize" > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_${TAG}.log
timeout 900 python tools/gpu_probe.py $SECS > gpurun_out/probe_${TAG}.log 2>&1
if [ "$3" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_minus -s 1 -c 1 \
      -o gpurun_out/prof_dgemm_big_${TAG} -f python tools/gpu_probe.py gemmbig > gpurun_out/ncu_dgemm_big_${TAG}.log 2>&1
fi
tail -4 gpurun_out/pytest_${TAG}.log
