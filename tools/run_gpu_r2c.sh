#!/bin/bash
mkdir -p gpurun_out
DBG_CFGS=${1:-260,261,262,263} timeout 300 python tools/dbg_batched.py > gpurun_out/r2c_dbg_batched.txt 2>&1
PROBE_CFGS=${2:-129,256,257,260,261,262,263} timeout 300 python tools/gpu_probe.py batched > gpurun_out/r2c_probe_batched.jsonl 2>&1
cat gpurun_out/r2c_dbg_batched.txt; python - <<'PY'
import json
for l in open('gpurun_out/r2c_probe_batched.jsonl'):
    d=json.loads(l)
    if 'bench' in d: print(d['bench'], d['cfg'], round(d['ms_best'],3), 'ms', round(d['mats_per_s']/1e6,1), 'M/s', round(d['frac_of_6453'],3))
PY
