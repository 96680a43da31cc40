#!/bin/bash
# Round-2 GPU check O (1 GPU): GPU suite + S bench after (Q|ja) columns travel with the poles in the host-pointer path.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02o_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02o_pytest_gpu.log
(time timeout 900 python bench.py --steps 1 --warmup 3) > gpurun_out/r02o_bench_S_short.json 2> gpurun_out/r02o_bench_S_short.err
tail -c 1200 gpurun_out/r02o_bench_S_short.json; tail -3 gpurun_out/r02o_bench_S_short.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
