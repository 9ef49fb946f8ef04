#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_multigpu.py "tests/test_gpu_parity.py::test_blocked_f32_matches_oracle" -m gpu -x -q --timeout=900 > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
( time timeout 1500 python bench.py --steps 2 --warmup 3 ) > gpurun_out/r2g_bench.log 2>&1
tail -5 gpurun_out/r2g_pytest.log; tail -c 6000 gpurun_out/r2g_bench.log
