#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
timeout 560 python bench.py > gpurun_out/bench_S_1gpu.json 2> gpurun_out/bench_S_1gpu.err; echo rc=$? wall=$(( $(date +%s) - T0 ))s
tail -c 600 gpurun_out/bench_S_1gpu.json; tail -n 2 gpurun_out/bench_S_1gpu.err
