#!/bin/bash
# Full evidence run: all GPU tests, smoke, bench lines, ncu launch list and full captures.  Usage: bash tools/run_gpu_final.sh <tag>
TAG=${1:-r1b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/pytest_full_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_full_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
bash tools/run_gpu_bench.sh ${TAG} > gpurun_out/run_bench_${TAG}.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_*_${TAG}.ncu-rep > gpurun_out/ncu_full_metrics_${TAG}.txt 2>&1
tail -4 gpurun_out/pytest_full_${TAG}.log; cat gpurun_out/smoke_${TAG}.log | tail -2
cat gpurun_out/bench_${TAG}.json gpurun_out/bench_c3_f64_${TAG}.json gpurun_out/bench_c3_f32_${TAG}.json gpurun_out/bench_ref_${TAG}.json
