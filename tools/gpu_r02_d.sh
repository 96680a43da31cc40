#!/bin/bash
# Round-2 GPU check D (2 GPUs): pole-sharded build -- multi-rank parity tests, a complete S-shape sharded build and a
# sampled one (the extrapolation path the L-shape run uses), against the replicated build of the same shape.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_multi_gpu.py -q -x) > gpurun_out/r02d_pytest_multi.log 2>&1
tail -15 gpurun_out/r02d_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(time timeout 600 $TR --master-port 29551 bench.py --gpus 2 --steps 1 --warmup 1 --sharded) > gpurun_out/r02d_bench_S_2gpu_sharded.json 2> gpurun_out/r02d_bench_S_2gpu_sharded.err
tail -c 2500 gpurun_out/r02d_bench_S_2gpu_sharded.json; grep -v "OpenBLAS\|^$" gpurun_out/r02d_bench_S_2gpu_sharded.err | tail -5
(time timeout 600 $TR --master-port 29552 bench.py --gpus 2 --steps 1 --warmup 1 --sharded --sample-limit 3) > gpurun_out/r02d_bench_S_2gpu_sampled.json 2> gpurun_out/r02d_bench_S_2gpu_sampled.err
tail -c 2500 gpurun_out/r02d_bench_S_2gpu_sampled.json; grep -v "OpenBLAS\|^$" gpurun_out/r02d_bench_S_2gpu_sampled.err | tail -5
