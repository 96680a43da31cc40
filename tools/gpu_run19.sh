#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python bench.py --shape 24,240,1000,264 --cpu-budget 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo rc=$?; tail -c 1500 gpurun_out/bench_small.json; tail -n 3 gpurun_out/bench_small.err
timeout 300 python tools/driver_timing.py 264,21,700 3 > gpurun_out/driver_timing2.log 2>&1; tail -n 5 gpurun_out/driver_timing2.log
