#!/bin/bash
mkdir -p gpurun_out
free -g | head -2; nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --shape 40,400,1600,440 --steps 2 --warmup 1 --cpu-budget 5 > gpurun_out/bench_medium.json 2> gpurun_out/bench_medium.err; echo rc=$?; cat gpurun_out/bench_medium.json; tail -3 gpurun_out/bench_medium.err
timeout 900 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_S_v0.json 2> gpurun_out/bench_S_v0.err; echo rc=$?; cat gpurun_out/bench_S_v0.json; tail -3 gpurun_out/bench_S_v0.err
