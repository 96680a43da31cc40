#!/bin/bash
# Round-2 GPU check N (1 GPU): GPU suite + S bench + nphys=367 after the stream-K cost-model change.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02n_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02n_pytest_gpu.log
(time timeout 900 python bench.py --steps 2 --warmup 3) > gpurun_out/r02n_bench_S_short.json 2> gpurun_out/r02n_bench_S_short.err
tail -c 1200 gpurun_out/r02n_bench_S_short.json; tail -3 gpurun_out/r02n_bench_S_short.err
for S in 100,1000,4000,367 1000,100,4000,367 64,100,4000,1100; do
  for OV in 1 0; do echo "== $S OVERLAP=$OV"; AGF2_B200_OVERLAP=$OV python tools/prof_case.py $S 2 2>&1 | grep -E "shape|rror" | tail -1; done
done 2>&1 | tee gpurun_out/r02n_p367.log
