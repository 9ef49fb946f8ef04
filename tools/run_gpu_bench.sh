#!/bin/bash
# Bench + ncu evidence run.  Usage: bash tools/run_gpu_bench.sh <tag>
# Numbers printed by the runs under ncu are never bench values; only bench_<tag>*.json are.
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --workload c3 --dtype f64 --steps 5 > gpurun_out/bench_c3_f64_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --workload c3 --dtype f32 --steps 5 > gpurun_out/bench_c3_f32_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
# every launch of ONE warm step (the profiler range bench.py --ncu-step opens) with its device time
# (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --ncu-step > gpurun_out/ncu_launches_${TAG}.log 2>&1
# top kernels, full sections, from the same step
for KS in dgemm_minus:4 panel_blocked_kernel:4 laswp_trsm:10 dtrsm_ll:0 laswp_kernel:4; do
  K=${KS%%:*}; SKIP=${KS##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${K} -s ${SKIP} -c 2 \
      -o gpurun_out/prof_${K}_${TAG} -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --ncu-step > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batched_lu32 -s 3 -c 1 \
    -o gpurun_out/prof_batched_f64_${TAG} -f python bench.py --workload c3 --dtype f64 --steps 1 --no-e2e --no-cpu > gpurun_out/ncu_batched_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batched_lu32 -s 3 -c 1 \
    -o gpurun_out/prof_batched_f32_${TAG} -f python bench.py --workload c3 --dtype f32 --steps 1 --no-e2e --no-cpu >> gpurun_out/ncu_batched_${TAG}.log 2>&1
cat gpurun_out/bench_${TAG}.json
tail -3 gpurun_out/bench_${TAG}.err
