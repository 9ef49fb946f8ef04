#!/bin/bash
# Verification run without profiler passes: all GPU tests, smoke, the bench lines.  Usage: bash tools/run_gpu_verify.sh <tag>
TAG=${1:-r1d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 > gpurun_out/pytest_full_${TAG}.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_full_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --workload c3 --dtype f64 --steps 5 > gpurun_out/bench_c3_f64_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --workload c3 --dtype f32 --steps 5 > gpurun_out/bench_c3_f32_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/pytest_full_${TAG}.log; tail -2 gpurun_out/smoke_${TAG}.log
cut -c1-260 gpurun_out/bench_${TAG}.json gpurun_out/bench_c3_f64_${TAG}.json gpurun_out/bench_c3_f32_${TAG}.json gpurun_out/bench_ref_${TAG}.json
tail -3 gpurun_out/bench_${TAG}.err
