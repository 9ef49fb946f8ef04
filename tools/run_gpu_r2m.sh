#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/r2m_pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2m_pytest_gpu.txt
tail -6 gpurun_out/r2m_pytest_gpu.txt
