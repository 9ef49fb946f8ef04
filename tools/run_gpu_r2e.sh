#!/bin/bash
mkdir -p gpurun_out
./tools/lsubench > gpurun_out/r2e_lsubench.jsonl 2>&1
bash tools/run_gpu_r2d.sh 262 r2e > /dev/null 2>&1
cat gpurun_out/ncu_batched_r2e_summary.txt
grep -E "shfl|redux" gpurun_out/r2e_lsubench.jsonl
