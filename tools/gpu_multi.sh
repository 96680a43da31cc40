#!/bin/bash
# usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_S_${N}gpu.json 2> gpurun_out/bench_S_${N}gpu.err; echo rc=$?
tail -1 gpurun_out/bench_S_${N}gpu.json | cut -c1-2500; tail -5 gpurun_out/bench_S_${N}gpu.err
