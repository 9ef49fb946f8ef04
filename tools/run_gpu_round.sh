#!/bin/bash
# First-contact GPU run: microbench, parity tests, probe.  Everything bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > gpurun_out/lscpu.txt 2>&1
timeout 120 tools/microbench > gpurun_out/microbench.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -k "not fullsize" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
echo "probe exit $?" >> gpurun_out/probe.log
tail -5 gpurun_out/pytest_gpu.log
