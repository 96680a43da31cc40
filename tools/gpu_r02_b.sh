#!/bin/bash
# Round-2 GPU check B (2 GPUs): multi-rank NCCL parity tests, the whole GPU suite, a short 2-GPU bench.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02b_gpu.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r02b_gpu.txt 2>&1
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r02b_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02b_pytest_gpu.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 3) > gpurun_out/r02b_bench_S_2gpu.json 2> gpurun_out/r02b_bench_S_2gpu.err
tail -c 2500 gpurun_out/r02b_bench_S_2gpu.json; grep -v "OpenBLAS\|^$" gpurun_out/r02b_bench_S_2gpu.err | tail -5
