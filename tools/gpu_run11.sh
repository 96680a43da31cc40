#!/bin/bash
# parity, small-shape bench sanity (all keys), then the full S bench and the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --shape 24,240,1000,264 --steps 2 --warmup 1 --cpu-budget 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err || { echo "small bench failed"; tail -20 gpurun_out/bench_small.err; exit 1; }
tail -c 3000 gpurun_out/bench_small.json
timeout 2400 python bench.py > gpurun_out/bench_S_r01b.json 2> gpurun_out/bench_S_r01b.err; echo rc=$?
tail -c 4000 gpurun_out/bench_S_r01b.json; tail -3 gpurun_out/bench_S_r01b.err
