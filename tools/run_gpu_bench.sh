#!/bin/bash
# Bench + ncu evidence run.  Usage: bash tools/run_gpu_bench.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --workload c3 --dtype f64 --steps 5 > gpurun_out/bench_c3_f64_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --workload c3 --dtype f32 --steps 5 > gpurun_out/bench_c3_f32_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
# every launch of one warm step with its device time (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 3200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
# top kernels, full sections
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_minus -s 40 -c 2 \
    -o gpurun_out/prof_dgemm_${TAG} -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_dgemm_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 40 -c 2 \
    -o gpurun_out/prof_panel_${TAG} -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_panel_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batched_lu32 -s 3 -c 1 \
    -o gpurun_out/prof_batched_f64_${TAG} -f python bench.py --workload c3 --dtype f64 --steps 1 --no-e2e --no-cpu > gpurun_out/ncu_batched_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batched_lu32 -s 3 -c 1 \
    -o gpurun_out/prof_batched_f32_${TAG} -f python bench.py --workload c3 --dtype f32 --steps 1 --no-e2e --no-cpu >> gpurun_out/ncu_batched_${TAG}.log 2>&1
cat gpurun_out/bench_${TAG}.json
tail -3 gpurun_out/bench_${TAG}.err
