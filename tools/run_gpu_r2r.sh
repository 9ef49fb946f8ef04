#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=600 -k "drain or chunked or blocked_f64_matches" > gpurun_out/r2r_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2r_pytest.log
tail -5 gpurun_out/r2r_pytest.log
timeout 900 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/r2r_bench_c2.json 2> gpurun_out/r2r_bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2r_bench_c2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','getrf_ms','e2e','e2e_getrf_pinned','e2e_getrf_pageable'):
    print(k, d.get(k))
PY
