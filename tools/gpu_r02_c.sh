#!/bin/bash
# Round-2 GPU check C (1 GPU): the driver's exact command lines, both arms, with wall times.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
S0=$(date +%s)
timeout 870 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02c_reference_1gpu.json 2> gpurun_out/r02c_reference_1gpu.err; echo "reference rc=$? wall=$(( $(date +%s) - S0 )) s" | tee gpurun_out/r02c_wall.txt
S1=$(date +%s)
timeout 870 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err; echo "b200 rc=$? wall=$(( $(date +%s) - S1 )) s" | tee -a gpurun_out/r02c_wall.txt
tail -c 1500 gpurun_out/r02c_bench_1gpu.json; tail -3 gpurun_out/r02c_bench_1gpu.err
