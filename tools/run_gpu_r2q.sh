#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/r2q_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2q_pytest.log
tail -6 gpurun_out/r2q_pytest.log
timeout 600 python tools/r2_probe_panel_push.py check phases coarse fine time getrf > gpurun_out/r2p_probe_panel_push.jsonl 2>&1
echo "probe exit $?" >> gpurun_out/r2p_probe_panel_push.jsonl
grep -E "check|getrf\"" gpurun_out/r2p_probe_panel_push.jsonl | tail -20
