#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_S_full.json 2> gpurun_out/bench_S_full.err; echo rc=$?; cat gpurun_out/bench_S_full.json; tail -3 gpurun_out/bench_S_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_S_ref.json 2> gpurun_out/bench_S_ref.err; cat gpurun_out/bench_S_ref.json | cut -c1-900
