#!/bin/bash
# Round-2 GPU check A (1 GPU): GPU test-suite, a short full-size bench of both arms with the wall-time breakdown.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02a_gpu.txt 2>&1
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r02a_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02a_pytest_gpu.log
(time timeout 900 python bench.py --steps 1 --warmup 2) > gpurun_out/r02a_bench_S_short.json 2> gpurun_out/r02a_bench_S_short.err
tail -c 3000 gpurun_out/r02a_bench_S_short.json; tail -3 gpurun_out/r02a_bench_S_short.err
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r02a_bench_S_reference.json 2> gpurun_out/r02a_bench_S_reference.err
tail -c 2500 gpurun_out/r02a_bench_S_reference.json; grep -v OpenBLAS gpurun_out/r02a_bench_S_reference.err | tail -3
