#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/r2a_box.txt
free -g >> gpurun_out/r2a_box.txt; nproc >> gpurun_out/r2a_box.txt
./tools/lsubench > gpurun_out/r2a_lsubench.jsonl 2>&1
timeout 600 python tools/r2_probe_first.py > gpurun_out/r2a_probe_first.jsonl 2>&1
PROBE_CFGS=-1 timeout 300 python tools/gpu_probe.py batched > gpurun_out/r2a_probe_batched.jsonl 2>&1
tail -5 gpurun_out/r2a_probe_first.jsonl; tail -3 gpurun_out/r2a_probe_batched.jsonl; head -70 gpurun_out/r2a_lsubench.jsonl
