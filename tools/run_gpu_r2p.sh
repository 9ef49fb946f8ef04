#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/r2_probe_panel_push.py check > gpurun_out/r2p_check.jsonl 2>&1
echo "exit $?" >> gpurun_out/r2p_check.jsonl
tail -8 gpurun_out/r2p_check.jsonl
timeout 600 python tools/r2_probe_panel_push.py phases coarse time getrf > gpurun_out/r2p_time.jsonl 2>&1
echo "exit $?" >> gpurun_out/r2p_time.jsonl
grep -E "phases|coarse|getrf" gpurun_out/r2p_time.jsonl | grep -v '"gen": 2' | tail -40
