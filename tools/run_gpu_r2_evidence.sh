#!/bin/bash
# Round-2 evidence run (one GPU): driver-style bench line, ncu launch list of the c2 step, ncu --set full of the dominant kernels.
# Numbers printed by the runs under ncu are never bench values.
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${TAG}_box.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${TAG}_box.txt; free -g | head -2 >> gpurun_out/${TAG}_box.txt
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_ncu_launch_list_c2_step.csv python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu --ncu-step > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dgemm_minus -s 7 -c 1 \
    -o gpurun_out/${TAG}_prof_dgemm -f python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu --ncu-step > gpurun_out/${TAG}_ncu_dgemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgemm_tf32x3 -s 1 -c 1 \
    -o gpurun_out/${TAG}_prof_sgemm_tf32x3 -f python tools/r2_sgemm_one.py > gpurun_out/${TAG}_ncu_sgemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_push -s 2 -c 1 \
    -o gpurun_out/${TAG}_prof_panel_push -f python tools/r2_panel_one.py d 8192 32 > gpurun_out/${TAG}_ncu_panel.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:trsm_strip -s 2 -c 1 \
    -o gpurun_out/${TAG}_prof_trsm_strip -f python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu --ncu-step > gpurun_out/${TAG}_ncu_trsm_strip.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batched_lu32 -s 3 -c 1 \
    -o gpurun_out/${TAG}_prof_batched_f64 -f python bench.py --workload c3 --dtype f64 --steps 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_batched.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batched_lu32 -s 3 -c 1 \
    -o gpurun_out/${TAG}_prof_batched_f32 -f python bench.py --workload c3 --dtype f32 --steps 3 --no-e2e --no-cpu >> gpurun_out/${TAG}_ncu_batched.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_prof_dgemm.ncu-rep gpurun_out/${TAG}_prof_sgemm_tf32x3.ncu-rep gpurun_out/${TAG}_prof_panel_push.ncu-rep gpurun_out/${TAG}_prof_trsm_strip.ncu-rep gpurun_out/${TAG}_prof_batched_f64.ncu-rep gpurun_out/${TAG}_prof_batched_f32.ncu-rep > gpurun_out/${TAG}_ncu_full_metrics.txt 2>&1
cat gpurun_out/${TAG}_ncu_full_metrics.txt
tail -3 gpurun_out/${TAG}_bench_n1.err
head -c 1500 gpurun_out/${TAG}_bench_n1.json
# final verification of the tree the evidence was taken from: the whole GPU suite and the smoke entry
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/${TAG}_pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.txt
tail -3 gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.txt 2>&1
tail -2 gpurun_out/${TAG}_smoke.txt
