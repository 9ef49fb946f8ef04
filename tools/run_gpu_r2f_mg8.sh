#!/bin/bash
# 8 GPUs: the driver's --gpus 8 bench line, parity of the multi-GPU LU against one GPU, the 8-GPU cases of the multi-GPU tests
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2f_bench_n8.json 2> gpurun_out/r2f_bench_n8.err
tail -2 gpurun_out/r2f_bench_n8.err; head -c 600 gpurun_out/r2f_bench_n8.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/mg_check.py > gpurun_out/r2f_mg_check_8gpu.log 2>&1
tail -4 gpurun_out/r2f_mg_check_8gpu.log
timeout 600 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q --timeout=500 > gpurun_out/r2f_pytest_multigpu_8gpu.txt 2>&1
tail -3 gpurun_out/r2f_pytest_multigpu_8gpu.txt
