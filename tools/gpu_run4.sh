#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --shape 40,400,1600,440 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_medium.json 2> gpurun_out/bench_medium.err; echo rc=$?; cat gpurun_out/bench_medium.json; tail -3 gpurun_out/bench_medium.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_S_v2.json 2> gpurun_out/bench_S_v2.err; echo rc=$?; cat gpurun_out/bench_S_v2.json; tail -3 gpurun_out/bench_S_v2.err
