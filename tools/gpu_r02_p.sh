#!/bin/bash
# Round-2 GPU check P (2 GPUs): the whole GPU suite incl. the multi-rank NCCL tests, and a short 2-GPU bench, final code.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r02p_pytest_gpu.log 2>&1
tail -6 gpurun_out/r02p_pytest_gpu.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 2 --warmup 3) > gpurun_out/r02p_bench_S_2gpu.json 2> gpurun_out/r02p_bench_S_2gpu.err
tail -c 1500 gpurun_out/r02p_bench_S_2gpu.json; grep -v "OpenBLAS\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02p_bench_S_2gpu.err | tail -5
