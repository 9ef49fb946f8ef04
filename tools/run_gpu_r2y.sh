#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=500 -k "paired or strip or drain or kernel_variants or blocked_f64_matches" > gpurun_out/r2y_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2y_pytest.log; tail -4 gpurun_out/r2y_pytest.log
timeout 900 python tools/r2_probe_pair.py > gpurun_out/r2y_probe_pair.jsonl 2>&1; cat gpurun_out/r2y_probe_pair.jsonl | tail -12
