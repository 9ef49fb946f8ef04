#!/bin/bash
# 2 GPUs: the multi-GPU tests (torchrun mg_check on 2 ranks, single-process batched entry) and the driver's --gpus 2 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q --timeout=600 > gpurun_out/r2f_pytest_multigpu_2gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2f_pytest_multigpu_2gpu.txt
tail -4 gpurun_out/r2f_pytest_multigpu_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
tail -2 gpurun_out/r2f_bench_n2.err; head -c 700 gpurun_out/r2f_bench_n2.json
