#!/bin/bash
mkdir -p gpurun_out
./tools/lsubench > gpurun_out/r2b_lsubench.jsonl 2>&1
